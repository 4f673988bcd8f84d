#!/usr/bin/env python
"""Minimal e24 training-step driver for ncu: W warm-up steps + N steps of the bench workload, nothing else.

    python tools/profile_step.py [--batch 64] [--warmup 3] [--steps 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--table", action="store_true", help="print every C-ABI call of the measured steps (CUDA-event times)")
    args = ap.parse_args()
    import torch
    import cpc_b200
    from cpc_b200 import configs
    dev = torch.device("cuda", 0)
    exp = configs.experiment("e24")
    tc = exp["training_config"]
    torch.manual_seed(0)
    model, pre, _ = configs.setup_model(exp["cqt_config"], exp["encoder_config"], exp["ar_model_config"], tc, device=dev)
    trainer = cpc_b200.ContrastiveEstimationTrainer(
        model=model, dataset=None, device=dev, regularization=tc["regularization"],
        score_over_all_timesteps=tc["score_over_all_timesteps"], score_function=tc["score_function"],
        preprocessing=pre, prediction_steps=tc["prediction_steps"], verbose=False)
    opt = cpc_b200.optim.Adam(model.parameters(), lr=tc["learning_rate"])
    model.train()
    g = torch.Generator().manual_seed(1234)
    x = (0.1 * torch.randn(args.batch, model.item_length, generator=g)).to(dev)
    prof = None
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            torch.cuda.synchronize()
            cpc_b200._lib.reset_launch_count()
            if args.table:
                prof = cpc_b200.ops.KernelProfiler()
                prof.__enter__()
        loss, _ = trainer.loss_on_batch(x)
        model.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    if prof is not None:
        prof.__exit__(None, None, None)
        total = 0.0
        for k in prof.summary():
            total += k["total_ms"] / args.steps
            print("%-105s n=%-3d %.4f ms/step  %.1f TFLOP/s  %.0f GB/s" % (k["key"], k["count"] // args.steps,
                                                                        k["total_ms"] / args.steps, k["tflops"], k["gbs"]))
        print("sum of own calls: %.3f ms/step" % total)
    print("loss %.5f, own kernel launches in the measured steps: %d" % (loss.item(), cpc_b200._lib.launch_count()))


if __name__ == "__main__":
    main()
