"""Diagnostic (GPU box): where do the e24 golden gradients of the CUDA path and of the CPU oracle part ways?

Builds experiments['e24'] on the GPU and OracleE24 on the CPU with the same seeded weights, runs the first golden batch
through both and compares, layer by layer of the AR model, the pre-activations (batch-norm outputs), the ReLU gate
patterns and the max-pool winners.  A gate whose pre-activation is within rounding distance of zero flips between any
two fp32 implementations and switches its whole downstream gradient on or off (see oracle/make_golden.py::golden_e24).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cpc_b200                                              # noqa: E402
import cpc_oracle as O                                       # noqa: E402
import cpc_oracle_model as OM                                # noqa: E402

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "e24_step.npz"), allow_pickle=False))
dev = torch.device("cuda:0")
exp = cpc_b200.configs.experiment("e24")
tc = exp["training_config"]
model, pre, _ = cpc_b200.configs.setup_model(exp["cqt_config"], exp["encoder_config"], exp["ar_model_config"], tc, device=dev)
OM.reseed_parameters(model.named_parameters())
oracle = OM.OracleE24(60, 16)
OM.reseed_parameters(oracle.named_parameters(), OM.oracle_e24_name_map())
batch, order = int(g["batch"]), [int(i) for i in g["order"]]
audio = OM.e24_audio(2 * batch, model.item_length, seed=int(g["audio_seed"]))[order[:batch]]
model.train(); oracle.train()

# GPU: hooks on the AR model's BatchNorm1d (pre-activations) and MaxPool1d inputs
gpu = {"bn": [], "pool_in": []}
hooks = []
for m in model.autoregressive_model.modules():
    if isinstance(m, nn.BatchNorm1d):
        hooks.append(m.register_forward_hook(lambda mod, i, o: gpu["bn"].append(o.detach().cpu())))
for blk in model.autoregressive_model.module_list:
    for m in blk.main_modules:
        if isinstance(m, nn.MaxPool1d):
            hooks.append(m.register_forward_hook(lambda mod, i, o: gpu["pool_in"].append(i[0].detach().cpu())))
enc_out = {}
hooks.append(model.encoder.register_forward_hook(lambda mod, i, o: enc_out.__setitem__("z", o.detach().cpu())))
with torch.no_grad():
    model(pre(audio.to(dev).unsqueeze(1)))
for h in hooks:
    h.remove()

# CPU oracle: same quantities
cpu = {"bn": [], "pool_in": []}
with torch.no_grad():
    x = O.preprocess(audio.unsqueeze(1), oracle.plan, phase=True)
    for i, b in enumerate(oracle.blocks):
        x = b(x)
        if i < 3:
            x = F.relu(x)
    z = x[:, :, 0, :]
    _, h = O.predictive_split(z, 60, 16)
    ar = oracle.ar
    for conv, bn, skip, pool in zip(ar.convs, ar.bns, ar.skips, ar.pooling):
        if pool > 1:
            cpu["pool_in"].append(h.clone())
        main = F.max_pool1d(h, pool, ceil_mode=True) if pool > 1 else h
        pre_act = bn(conv(main))
        cpu["bn"].append(pre_act.clone())
        main = F.relu(pre_act)
        r = F.max_pool1d(h, pool, ceil_mode=True) if pool > 1 else h
        h = main + skip(r)[:, :, -main.shape[2]:]
print("encoder output: rel diff %.2e" % float((enc_out["z"] - z).norm() / z.norm()))
for l, (a, b) in enumerate(zip(gpu["bn"], cpu["bn"])):
    flips = (a > 0) != (b > 0)
    idx = flips.nonzero()
    print("AR layer %d: %d gates, max |diff| %.2e, rel diff %.2e, flipped gates %d %s" % (
        l, a.numel(), float((a - b).abs().max()), float((a - b).norm() / b.norm()), int(flips.sum()),
        [(tuple(i.tolist()), float(a[tuple(i)]), float(b[tuple(i)])) for i in idx[:4]]))
    near = int((b.abs() < 1e-5).sum())
    print("            gates of the oracle within 1e-5 of zero: %d (smallest %.2e)" % (near, float(b.abs().min())))
for l, (a, b) in enumerate(zip(gpu["pool_in"], cpu["pool_in"])):
    ia = F.max_pool1d(a, 2, ceil_mode=True, return_indices=True)[1]
    ib = F.max_pool1d(b, 2, ceil_mode=True, return_indices=True)[1]
    print("AR pool %d: winners differ in %d of %d windows" % (l, int((ia != ib).sum()), ia.numel()))
