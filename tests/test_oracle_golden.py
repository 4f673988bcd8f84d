"""CPU: pins oracle/cpc_oracle.py against the golden vectors the UNMODIFIED reference produced
(oracle/make_golden.py).  The GPU parity tests then compare the CUDA path with this oracle."""
import json
import math
import random

import numpy as np
import pytest
import torch

import cpc_oracle as O
from conftest import load_golden, rel_err


@pytest.fixture(scope="module")
def plan():
    return O.CqtPlan(16000, 30, 256, 32, 0.5, 128)


def test_filterbank_structural_anchors(plan):
    # the only facts the reference exposes about librosa's filterbank
    assert plan.kernel_sizes == [16384, 8192, 4096, 2048, 1024, 512, 256, 128, 64]
    assert plan.ranges == [(0, 19), (19, 51), (51, 83), (83, 115), (115, 147), (147, 179), (179, 211), (211, 243),
                           (243, 256)]
    assert abs(plan.lengths.sum() - 566110.2) < 0.1
    assert abs(plan.lengths[0] - 12178.146) < 1e-3 and abs(plan.lengths[-1] - 48.6125) < 1e-4
    # L1 normalisation and centring
    assert np.allclose(np.abs(plan.bank).sum(axis=1), 1.0)


def test_cqt_matches_reference(plan):
    g = load_golden("cqt.npz")
    x = torch.from_numpy(g["x"])
    assert rel_err(O.cqt_forward(x, plan), g["complex"]) < 1e-6
    assert rel_err(O.preprocess(x, plan, phase=False), g["logpow"]) < 1e-6
    y = O.preprocess(x, plan, phase=True)
    assert rel_err(y[:, 0], g["logpow_phase"][:, 0]) < 1e-6
    assert float((y[:, 1] - torch.from_numpy(g["logpow_phase"][:, 1])).abs().max()) < 1e-4
    assert rel_err(O.preprocess(x, plan, offset_zero=True, output_power=2., pooling=[1, 2], scaling=10.),
                   g["offset_pool_power"]) < 1e-6
    plan2 = O.CqtPlan(8000, 55, 120, 24, 1.0, 64)
    assert plan2.kernel_sizes == list(g["kernel_sizes2"])
    assert rel_err(O.cqt_forward(torch.from_numpy(g["x2"]), plan2), g["complex2"]) < 1e-6


def test_high_res_cqt_matches_reference():
    """cqt_high_res_dict (configs/cqt_configs.py:9-12; experiments e27 ... e32): 11 octave groups, 65 536 ... 64 taps."""
    g = load_golden("cqt_high_res.npz")
    cfg = json.loads(str(g["cfg"]))
    plan = O.CqtPlan(cfg["sample_rate"], cfg["fmin"], cfg["n_bins"], cfg["bins_per_octave"], cfg["filter_scale"], cfg["hop_length"])
    assert plan.kernel_sizes == list(g["kernel_sizes"]) and [list(r) for r in plan.ranges] == g["ranges"].tolist()
    x = torch.from_numpy(g["x"])
    assert rel_err(O.cqt_forward(x, plan), g["complex"]) < 1e-6
    y = O.preprocess(x, plan, phase=True)
    assert rel_err(y[:, 0], g["phase"][:, 0]) < 1e-6
    assert rel_err(O.preprocess(x, plan, offset_zero=True, pooling=[1, 2]), g["offset_pool"]) < 1e-6


def test_cqt_input_gradients_match_reference():
    """d/d(audio) through the oracle's CQT and phase scalogram against the reference's autograd (cqt_grad.npz); also the
    PhaseAccumulation formula (constant_q_transform.py:306-313) restated inline."""
    g = load_golden("cqt_grad.npz")
    plan = O.CqtPlan(8000, 55, 96, 24, 0.5, 64)
    assert plan.kernel_sizes == list(g["kernel_sizes"])
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    z = O.cqt_forward(x, plan)
    assert rel_err(z, g["z"]) < 1e-6
    (z * torch.from_numpy(g["gz"])).sum().backward()
    assert rel_err(x.grad, g["gx"]) < 1e-5
    x2 = torch.from_numpy(g["x2"]).requires_grad_(True)
    y = O.preprocess(x2, plan, phase=True, offset_zero=True, output_power=1., pooling=[1, 2], scaling=3.)
    assert rel_err(y[:, 0], g["y2"][:, 0]) < 1e-6
    (y * torch.from_numpy(g["gy2"])).sum().backward()
    assert rel_err(x2.grad, g["gx2"]) < 1e-3                      # 1/|z| factors: fp32 summation order matters
    fixed, scaling = O.phase_constants(plan)
    ph = torch.from_numpy(g["acc_in"])
    acc = torch.cumsum(torch.cat([torch.zeros(1, 96, 1), ph / torch.from_numpy(scaling).view(1, -1, 1)
                                  - torch.from_numpy(fixed).view(1, -1, 1)], dim=2), dim=2) % (2 * np.pi) - np.pi
    d = torch.remainder(acc - torch.from_numpy(g["acc_out"]) + np.pi, 2 * np.pi) - np.pi
    assert float(d.abs().max()) < 1e-4


def test_audio_encoder_matches_reference():
    g = load_golden("audio_encoder.npz")
    assert list(g["kat_shape"]) == [7, 32, 28]                 # tests/test_audioEncoder.py:19-25
    assert list(g["kat_rf_ds"]) == [465, 160]                  # tests/test_audioEncoder.py:27-48
    assert O.audio_encoder_geometry([10, 8, 4, 4, 4], [5, 4, 2, 2, 2]) == (465, 160)
    ws = [torch.from_numpy(g["p.layers.%d.weight" % i]).requires_grad_(True) for i in range(5)]
    bs = [torch.from_numpy(g["p.layers.%d.bias" % i]).requires_grad_(True) for i in range(5)]
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = O.audio_encoder_forward(x, ws, bs, [5, 4, 2, 2, 2])
    assert rel_err(y, g["y"]) < 1e-6
    (y * torch.from_numpy(g["gy"])).sum().backward()
    assert rel_err(x.grad, g["gx"]) < 1e-5
    for i in range(5):
        assert rel_err(ws[i].grad, g["g.layers.%d.weight" % i]) < 1e-5


def small_resnet_blocks():
    base = {'hidden_channels': None, 'kernel_size_1': (3, 3), 'kernel_size_2': (3, 3), 'top_padding_1': None,
            'top_padding_2': None, 'padding_1': 0, 'padding_2': 0, 'stride_1': 1, 'stride_2': 1, 'pooling_1': 1,
            'pooling_2': 1, 'bias': True, 'separable': False, 'residual': True, 'batch_norm': False,
            'ceil_pooling': False}
    b0 = dict(base, in_channels=2, out_channels=8, kernel_size_2=(9, 1), top_padding_2=8, stride_1=2, batch_norm=True)
    b1 = dict(base, in_channels=8, out_channels=16, kernel_size_2=(6, 1), stride_1=2, batch_norm=True, padding_1=1)
    b2 = dict(base, in_channels=16, out_channels=24, kernel_size_1=(2, 2), kernel_size_2=(1, 1), pooling_1=2,
              ceil_pooling=True)
    return [b0, b1, b2]


def block_param_map(g, i, cfg, prefix="p."):
    """reference state_dict keys (main_modules.N...) -> the oracle's functional names."""
    keys = sorted(k for k in g if k.startswith("%sblocks.%d." % (prefix, i)))
    convs = sorted({int(k.split(".")[4]) for k in keys if ".main_modules." in k and k.endswith(".weight")
                    and g[k].ndim == 4})
    bns = sorted({int(k.split(".")[4]) for k in keys if ".main_modules." in k and k.endswith("running_mean")})
    p = {}
    for name, idx in zip(("conv_a", "conv_b"), convs):
        for leaf in ("weight", "bias"):
            k = "%sblocks.%d.main_modules.%d.%s" % (prefix, i, idx, leaf)
            if k in g:
                p[name + "." + leaf] = torch.from_numpy(g[k])
    for name, idx in zip(("bn_a", "bn_b"), bns):
        for leaf in ("weight", "bias"):
            p[name + "." + leaf] = torch.from_numpy(g["%sblocks.%d.main_modules.%d.%s" % (prefix, i, idx, leaf)])
    for k in keys:
        if ".residual_modules." in k and k.endswith(".weight"):
            p["res.weight"] = torch.from_numpy(g[k])
    return p


def test_resnet_encoder_matches_reference():
    g = load_golden("resnet_encoder.npz")
    cfgs = small_resnet_blocks()
    params = [block_param_map(g, i, c) for i, c in enumerate(cfgs)]
    x = torch.from_numpy(g["x"])
    y = O.residual_encoder_forward(x, cfgs, params, training=True)
    assert tuple(y.shape) == g["y"].shape
    assert rel_err(y, g["y"]) < 1e-5


@pytest.mark.parametrize("tag,all_steps,kind", [("a", True, "linear"), ("p", False, "softplus")])
def test_gradient_penalty_step_matches_reference_trainer(tag, all_steps, kind):
    """First training step of the reference with wasserstein_gradient_penalty=True (trainer_gp.npz): the oracle's
    CQT -> residual encoder -> conv AR -> W_k -> InfoNCE + gradient penalty, evaluated on the reference's initial
    parameters and its first batch, reproduces the logged loss and max score; one SGD step on the oracle's gradients
    reproduces the reference's parameters after that step."""
    import torch.nn.functional as F
    full = load_golden("trainer_gp.npz")
    g = {k[len(tag) + 1:]: v for k, v in full.items() if k.startswith(tag + ".")}
    lr, bs = float(g["lr"]), int(g["batch_size"])
    random.seed(5)
    first = O.file_batch_sampler([g["items"].shape[0]], bs)[0]
    audio = torch.from_numpy(g["items"][first]).unsqueeze(1)
    cfgs = small_resnet_blocks()
    cfgs[2] = dict(cfgs[2], kernel_size_1=(30, 2), pooling_1=1, ceil_pooling=False)
    cfgs[1] = dict(cfgs[1], kernel_size_2=(35, 1))
    sd = {"p." + k[3:]: v for k, v in g.items() if k.startswith("s0.")}
    enc_state = {k.replace("p.encoder.", "p."): v for k, v in sd.items() if k.startswith("p.encoder.")}
    params = [block_param_map(enc_state, i, c) for i, c in enumerate(cfgs)]
    leaves = {}

    def leaf(name, key):
        t = torch.from_numpy(sd["p." + key]).clone().requires_grad_(True)
        leaves[key] = t
        return t
    for i, p in enumerate(params):                                # make every encoder parameter a leaf we can read back
        keys = sorted(k for k in enc_state if k.startswith("p.blocks.%d." % i))
        convs = sorted({int(k.split(".")[4]) for k in keys if ".main_modules." in k and k.endswith(".weight")
                        and enc_state[k].ndim == 4})
        bns = sorted({int(k.split(".")[4]) for k in keys if ".main_modules." in k and k.endswith("running_mean")})
        names = {}
        for name, idx in list(zip(("conv_a", "conv_b"), convs)) + list(zip(("bn_a", "bn_b"), bns)):
            for leaf_name in ("weight", "bias"):
                names[name + "." + leaf_name] = "blocks.%d.main_modules.%d.%s" % (i, idx, leaf_name)
        for k in keys:
            if ".residual_modules." in k and k.endswith(".weight"):
                names["res.weight"] = k[2:]
        for name in list(p):
            p[name] = leaf(name, "encoder." + names[name])
    ar = "autoregressive_model.module_list."
    w0, b0 = leaf("w0", ar + "0.main_modules.0.weight"), leaf("b0", ar + "0.main_modules.0.bias")
    g0, h0 = leaf("g0", ar + "0.main_modules.1.weight"), leaf("h0", ar + "0.main_modules.1.bias")
    w1, b1 = leaf("w1", ar + "1.main_modules.1.weight"), leaf("b1", ar + "1.main_modules.1.bias")
    g1, h1 = leaf("g1", ar + "1.main_modules.2.weight"), leaf("h1", ar + "1.main_modules.2.bias")
    wk = leaf("wk", "prediction_model.weight")
    plan = O.CqtPlan(16000, 30, 256, 32, 0.5, 128)
    with torch.no_grad():
        scal = O.preprocess(audio, plan, phase=True)
    scal.requires_grad_(True)
    z = O.residual_encoder_forward(scal, cfgs, params, training=True)
    targets, vis = O.predictive_split(z, 10, 3)
    x = F.relu(F.batch_norm(F.conv1d(vis, w0, b0), None, None, g0, h0, training=True))
    x = F.max_pool1d(x, 2, ceil_mode=True)
    x = F.relu(F.batch_norm(F.conv1d(x, w1, b1), None, None, g1, h1, training=True))
    pred = F.linear(x[:, :, -1], wk).view(-1, 3, 24)
    loss, mx = O.infonce_loss(pred, targets, all_steps, kind, 0.25)
    loss = loss + O.gradient_penalty(pred, targets, scal, all_steps, kind, 10.0)
    assert abs(float(loss.detach()) - g["losses"][0]) < 1e-4 * abs(g["losses"][0]), (float(loss.detach()), g["losses"][0])
    assert abs(float(mx.detach()) - g["max_scores"][0]) < 1e-4 * max(1.0, abs(g["max_scores"][0]))
    loss.backward()
    from conftest import bn_shadowed_biases, grad_err
    noise_only = bn_shadowed_biases([k[3:] for k in g if k.startswith("s0.")])
    checked = 0
    for key, t in leaves.items():
        if key in noise_only:
            continue
        ref_grad = (torch.from_numpy(g["s0." + key]) - torch.from_numpy(g["s1." + key])) / lr
        # (before - after) / lr with lr = 1e-4: parameters near 1.0 resolve 6e-8, i.e. 6e-4 absolute on a gradient
        assert grad_err(t.grad, ref_grad) < 5e-3, key
        checked += 1
    assert checked >= 20


def test_infonce_matches_reference_trainer():
    g = load_golden("infonce.npz")
    cases = json.loads(str(g["cases"]))
    assert len(cases) == 24
    for c in cases:
        t = c["tag"]
        pred, tgt = torch.from_numpy(g[t + ".pred"]), torch.from_numpy(g[t + ".tgt"])
        loss, mx, dp, dz = O.infonce_with_grads(pred, tgt, c["all_steps"], c["kind"], c["reg"])
        assert abs(float(loss) - float(g[t + ".loss"])) < 2e-6 * max(1.0, abs(float(loss))), c
        assert abs(float(mx) - float(g[t + ".max"])) < 1e-5, c
        assert rel_err(dp, g[t + ".dpred"]) < 1e-4, c          # golden grads went through an fp32 SGD update
        assert rel_err(dz, g[t + ".dtgt"]) < 1e-4, c
        clean = O.infonce_loss_clean(pred.double(), tgt.double(), c["all_steps"], c["kind"], c["reg"])
        assert abs(float(clean) - float(loss)) < 1e-9, c       # the scramble is loss-neutral


def test_sampler_bit_exact():
    for c in load_golden("sampler.json"):
        if c["global_seed"] is not None:
            random.seed(c["global_seed"])
        # the reference mutates per-file index lists in place across epochs -> replay on persistent lists
        epochs = []
        if c["file_batch_size"] == 1:
            for _ in range(2):
                epochs.append(O.file_batch_sampler(c["counts"], c["batch_size"], 1, True, c["seed"]))
            assert epochs == c["epochs"], c
        else:
            first = O.file_batch_sampler(c["counts"], c["batch_size"], c["file_batch_size"], True, c["seed"])
            assert first == c["epochs"][0], c


def test_validation_metrics_match_reference_validate():
    """oracle.validation_metrics vs the reference's own validate() (contrastive_estimation_training.py:178-269)."""
    g = load_golden("validate.npz")
    for c in json.loads(str(g["cases"])):
        t = c["tag"]
        pred, tgt = torch.from_numpy(g[t + ".pred"]), torch.from_numpy(g[t + ".tgt"])
        losses, acc, score = O.validation_metrics(pred, tgt, c["all_steps"], c["kind"])
        assert torch.allclose(losses, torch.from_numpy(g[t + ".losses"]), rtol=1e-5, atol=1e-5), c
        assert torch.equal(acc, torch.from_numpy(g[t + ".acc"])), c
        assert abs(float(score) - float(g[t + ".score"])) < 1e-5 * max(1.0, abs(float(g[t + ".score"]))), c
        n = c["b"] * c["k"] if c["all_steps"] else c["b"]
        assert torch.allclose(math.log(n) - losses, torch.from_numpy(g[t + ".mi"]), rtol=1e-5, atol=1e-5), c


def check_e24_against_golden(g, names, losses, max_scores, z, grads, params_after, tol, update_tol):
    """Shared by the CPU oracle pin below and the GPU parity test: ``names`` are the reference's state_dict keys,
    ``grads`` / ``params_after`` map them to full tensors (step-1 gradients, parameters after the second Adam step).

    Forward quantities are held to ``tol``.  Each gradient is held to max(tol, 3 x the reference's own self-noise):
    make_golden.py measured how far the reference's gradients move when its scalogram changes by 1e-7 / 1e-6 relative
    (about one fp32 ulp of the log-power values) or every conv / linear output by 2e-6 / 5e-6 relative -- ReLU gates near zero
    make them move by up to a few 1e-2 (``sn6.*``, ``sn7.*``, ``snl.*``, ``snl5.*`` in the fixture), so no independent implementation
    can be closer than that to one particular run of the reference."""
    import cpc_oracle_model as OM
    assert abs(losses[0] - float(g["losses"][0])) < tol * abs(float(g["losses"][0])), (losses, g["losses"])
    assert abs(max_scores[0] - float(g["max_scores"][0])) < tol * abs(float(g["max_scores"][0])), (max_scores, g["max_scores"])
    assert rel_err(z, g["z"]) < tol
    shadowed = set(json.loads(str(g["bn_shadowed"])))
    report = {}
    for n in names:
        idx = OM.subsample_index(n, grads[n].numel())
        ref = torch.from_numpy(g["g." + n]).double()
        mine = grads[n].detach().reshape(-1).cpu()[idx].double()
        if n in shadowed:
            # conv bias in front of a train-mode batch norm: the true gradient is exactly zero, both sides hold rounding noise
            assert float(mine.abs().max()) < 1e-3 * max(1.0, float(ref.abs().max()) * 1e3), n
            continue
        err = float((mine - ref).norm() / ref.norm().clamp_min(1e-30))
        bound = max(tol, 3.0 * max(float(g[k + n]) for k in ("sn6.", "sn7.", "snl.", "snl5.")))
        full = float(grads[n].detach().double().norm())
        report[n] = (max(err, abs(full - float(g["gn." + n])) / float(g["gn." + n])), bound)
    failed = {n: v for n, v in report.items() if not v[0] < v[1]}
    assert not failed, "gradients outside max(tol, 3 x reference self-noise): %s" % failed
    # second step: Adam has moved every weight by ~lr * sign(gradient); the second loss and the parameters see it
    assert abs(losses[1] - float(g["losses"][1])) < update_tol * abs(float(g["losses"][1])), (losses, g["losses"])
    lr = float(g["lr"])
    moved_wrong = 0.0
    total = 0
    for n in names:
        idx = OM.subsample_index(n, params_after[n].numel())
        ref = torch.from_numpy(g["p2." + n]).double()
        mine = params_after[n].detach().reshape(-1).cpu()[idx].double()
        assert float((mine - ref).abs().max()) < 4.5 * lr, n           # two Adam steps move an entry by at most ~2 lr each side
        if n not in shadowed:
            moved_wrong += float(((mine - ref).abs() > 0.25 * lr).double().sum())
            total += len(idx)
    # entries whose two-step Adam update disagrees by > lr/4.  Adam's first steps move an entry by ~lr * sign(gradient), so
    # this counts gradient entries whose sign is within the gradient noise discussed above (oracle vs reference: 0.9 %)
    report["__update_mismatch_fraction"] = (moved_wrong / total, 0.2)
    assert moved_wrong / total < 0.2
    return report


def test_oracle_e24_full_size_step_matches_reference():
    """BASELINE configs[1] at full item length (L = 97 024, batch 4): OracleE24 -- the CPU arm bench.py times --
    against two training steps of the reference's setup_model(experiments['e24']) + train() (e24_step.npz)."""
    import cpc_oracle_model as OM
    g = load_golden("e24_step.npz")
    names = json.loads(str(g["names"]))
    torch.manual_seed(0)
    model = OM.OracleE24(int(g["visible_steps"]), int(g["prediction_steps"]))
    assert model.item_length == int(g["item_length"])
    name_map = OM.oracle_e24_name_map()
    assert sorted(name_map) == sorted(names)
    OM.reseed_parameters(model.named_parameters(), name_map)
    own = dict(model.named_parameters())
    batch, order = int(g["batch"]), [int(i) for i in g["order"]]
    audio = OM.e24_audio(2 * batch, model.item_length)
    assert np.array_equal(audio[:, ::4099].numpy(), g["audio_check"])
    assert str(g["score_kind"]) == "linear_score_function" and bool(g["all_steps"]) and float(g["regularization"]) == 0.0
    opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]))
    model.train()
    losses, maxes, grads, z1 = [], [], None, None
    for step in range(2):
        pred, targets = model(audio[order[step * batch:(step + 1) * batch]])
        loss, mx = O.infonce_loss(pred, targets, True, "linear", 0.0)
        model.zero_grad()
        loss.backward()
        if step == 0:
            grads = {ref: own[o].grad.detach().clone() for ref, o in name_map.items()}
            z1 = model.last_z.clone()
        opt.step()
        losses.append(float(loss.detach()))
        maxes.append(float(mx))
    report = check_e24_against_golden(g, names, losses, maxes, z1, grads, {ref: own[o] for ref, o in name_map.items()},
                                      tol=1e-4, update_tol=2e-2)
    print("oracle e24: losses", losses, "worst gradient error",
          max(((k, v) for k, v in report.items() if not k.startswith("__")), key=lambda kv: kv[1][0]),
          "update mismatch fraction", report["__update_mismatch_fraction"][0])
