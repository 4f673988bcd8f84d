"""Worker of tests/test_gpu_parity.py::test_two_rank_nccl_parity (launched with torch.distributed.run, one rank per GPU).

SURVEY 8e: every rank trains on its contiguous shard of the global batch with its own negatives, gradients are averaged.
  1. raw-wave model: per-rank loss and gradients against the CPU oracle evaluated on THAT shard; the gradients the
     overlapped in-graph all-reduce leaves in the flat buffer against the mean of the per-shard ORACLE gradients.
  2. e24 at the full item length: per-rank loss against the oracle on that shard, and the all-reduced gradients against
     the mean of the per-rank gradients of the same kernels (plumbing exactness, free of ReLU-gate noise).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import cpc_b200                                              # noqa: E402
import cpc_oracle as O                                       # noqa: E402
import cpc_oracle_model as OM                                # noqa: E402
from cpc_b200 import ddp                                     # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gather_mean(t, world):
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return sum(parts) / world


def main():
    rank, world, local = ddp.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    assert world >= 2

    # ---- 1. raw-wave model against the oracle, shard by shard ----------------------------------------------------
    chans = [16, 24, 24, 24, 32]
    torch.manual_seed(0)
    enc = cpc_b200.AudioEncoder({'strides': [5, 4, 2, 2, 2], 'kernel_sizes': [10, 8, 4, 4, 4], 'channel_count': chans,
                                 'bias': True})
    model = cpc_b200.AudioPredictiveCodingModel(enc, cpc_b200.AudioGRUModel(32, 16), enc_size=32, ar_size=16,
                                                visible_steps=9, prediction_steps=4).to(dev)
    ddp.broadcast_parameters(model, 0)
    per = 6
    g = torch.Generator().manual_seed(77)
    global_batch = 0.1 * torch.randn(per * world, model.item_length, generator=g)
    shard = ddp.shard_batch(global_batch, rank, world)
    assert torch.equal(shard, global_batch[rank * per:(rank + 1) * per])
    trainer = cpc_b200.ContrastiveEstimationTrainer(model=model, dataset=None, device=dev, regularization=1.0,
                                                    score_over_all_timesteps=False,
                                                    score_function=cpc_b200.softplus_score_function, prediction_steps=4,
                                                    verbose=False)
    assert (trainer.rank, trainer.world) == (rank, world)
    oracle = OM.OracleRawWave(9, 4, channels=chans, ar_size=16)
    own = dict(model.named_parameters())
    name_map = {}
    for i in range(5):
        name_map["encoder.layers.%d.weight" % i], name_map["encoder.layers.%d.bias" % i] = "convs.%d.weight" % i, "convs.%d.bias" % i
    for leaf in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
        name_map["autoregressive_model.gruCell." + leaf] = "gru." + leaf
    name_map["prediction_model.weight"] = "predict.weight"
    theirs = dict(oracle.named_parameters())
    assert set(name_map) == set(own) and set(name_map.values()) == set(theirs)
    with torch.no_grad():
        for k, v in name_map.items():
            theirs[v].copy_(own[k].detach().cpu())
    pred, tgt = oracle(shard)
    want_loss, _ = O.infonce_loss(pred, tgt, False, "softplus", 1.0)
    want_loss.backward()
    # graphed multi-rank step with SGD(lr=0): weights stay, the flat buffer keeps the summed gradients
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    step = cpc_b200.GraphedTrainStep(trainer, opt, (per, model.item_length), warmup=2)
    assert "captured inside" in step.overlap_description, step.overlap_description
    loss, _ = step(shard.to(dev))
    torch.cuda.synchronize()
    assert abs(loss.item() - float(want_loss)) < 1e-3 * abs(float(want_loss)), (loss.item(), float(want_loss))
    for k, v in name_map.items():
        mean_oracle = gather_mean(theirs[v].grad.to(dev), world)
        mine = own[k].grad                                      # a view of the flat buffer: the averaged gradient
        assert rel(mine, mean_oracle) < 3e-3, (k, rel(mine, mean_oracle))
    step.remove_hooks()

    # ---- 2. e24 at full item length: per-shard loss vs oracle; all-reduced gradients vs mean of per-rank gradients --------
    exp = cpc_b200.configs.experiment("e24")
    tc = exp["training_config"]
    torch.manual_seed(0)
    model, pre, _ = cpc_b200.configs.setup_model(exp["cqt_config"], exp["encoder_config"], exp["ar_model_config"], tc, device=dev)
    OM.reseed_parameters(model.named_parameters())
    per = 2
    audio = OM.e24_audio(per * world, model.item_length, seed=4321)
    shard = ddp.shard_batch(audio, rank, world)
    trainer = cpc_b200.ContrastiveEstimationTrainer(model=model, dataset=None, device=dev, regularization=tc["regularization"],
                                                    score_over_all_timesteps=tc["score_over_all_timesteps"],
                                                    score_function=tc["score_function"], preprocessing=pre,
                                                    prediction_steps=tc["prediction_steps"], verbose=False)
    model.train()
    oracle = OM.OracleE24(60, 16)
    OM.reseed_parameters(oracle.named_parameters(), OM.oracle_e24_name_map())
    oracle.train()
    with torch.no_grad():
        pred, tgt = oracle(shard)
        want_loss, _ = O.infonce_loss(pred, tgt, True, "linear", 0.0)
    # per-rank gradients of the same kernels, eagerly, without any collective
    bn_state = {k: v.clone() for k, v in model.state_dict().items()}
    # (on a side stream, and without keeping the loss tensor: a live autograd graph would pin the parameters' gradient
    # accumulators to the stream of THIS pass -- the legacy default stream cannot take part in a later graph capture)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        loss_t, _ = trainer.loss_on_batch(shard.to(dev))
        model.zero_grad(set_to_none=True)
        loss_t.backward()
        loss_value = loss_t.item()
        del loss_t
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    assert abs(loss_value - float(want_loss)) < 1e-3 * abs(float(want_loss)), (loss_value, float(want_loss))
    local = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    model.load_state_dict(bn_state)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    step = cpc_b200.GraphedTrainStep(trainer, opt, (per, model.item_length), warmup=2)
    assert "captured inside" in step.overlap_description, step.overlap_description
    loss2, _ = step(shard.to(dev))
    torch.cuda.synchronize()
    assert abs(loss2.item() - loss_value) < 1e-5 * abs(loss_value)
    worst = 0.0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import bn_shadowed_biases
    noise_only = bn_shadowed_biases(model.state_dict().keys())   # true gradient exactly 0: rounding noise on both sides
    for n, p in model.named_parameters():
        want = gather_mean(local[n], world)
        if n in noise_only:
            continue
        err = float((p.grad.double() - want.double()).norm() / want.double().norm().clamp_min(1e-12))
        worst = max(worst, err)
        assert err < 2e-4, (n, err)                               # same kernels; only the order of fp32 atomics differs
    dist.barrier()
    print("RANK_OK %d (e24 shard loss %.6f vs oracle %.6f, worst all-reduce deviation %.1e, %s)"
          % (rank, loss_value, float(want_loss), worst, step.overlap_description))


if __name__ == "__main__":
    try:
        main()
    except BaseException:                                         # noqa: BLE001 -- report compactly, never hang the launcher
        import traceback
        print("RANK_FAILED %s\n%s" % (os.environ.get("RANK"), traceback.format_exc()[-2500:]), flush=True)
        os._exit(1)
    os._exit(0)                                                   # skip interpreter teardown of graphs holding NCCL work
