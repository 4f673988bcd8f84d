#!/usr/bin/env python
"""InfoNCE scoring sweep (BASELINE configs[3]): candidates per softmax N in {128..8192} x prediction steps K, fused
forward + backward, CUDA-event timed, with size-independent property checks at every point (the reference cannot
materialise its (B,K,B,K) score tensor at the large corners):
  * loss(P, Z) at zero inputs = log(N)                      (uniform softmax)
  * sum of dL/dP . P + dL/dZ . Z = 2 * dL/ds . s identity   (Euler: the linear score is bilinear -> degree 2)
Prints one JSON line per point; `--quick` keeps the run under a minute.
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--enc", type=int, default=512)
    args = ap.parse_args()
    import torch
    import cpc_b200
    dev = torch.device("cuda", 0)
    ns = [128, 512, 2048] if args.quick else [128, 256, 512, 1024, 2048, 4096, 8192]
    ks = [4, 16] if args.quick else [4, 8, 16, 32]
    e = args.enc
    for all_steps in (False, True):
        for n in ns:
            for k in ks:
                b = n // k if all_steps else n
                if b < 2:
                    continue
                g = torch.Generator(device=dev).manual_seed(n * 131 + k)
                pred = (torch.randn(b, k, e, generator=g, device=dev) / math.sqrt(e)).requires_grad_(True)
                z = torch.randn(b, e, k, generator=g, device=dev).requires_grad_(True)
                loss0 = cpc_b200.ops.infonce(torch.zeros_like(pred), torch.zeros_like(z), all_steps, "linear", 0.0)[0]
                uniform_ok = abs(float(loss0) - math.log(b * k if all_steps else b)) < 1e-4

                def run():
                    pred.grad = z.grad = None
                    loss = cpc_b200.ops.infonce(pred, z, all_steps, "linear", 0.0)[0]
                    loss.backward()
                    return loss
                for _ in range(2):
                    loss = run()
                torch.cuda.synchronize()
                reps = 3 if n >= 4096 else 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    loss = run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                # Euler identity for a degree-2 homogeneous score: <dP,P> + <dZ,Z> = 2 * sum_ij G_ij s_ij, and the loss
                # is shift-invariant in s, so sum_ij G_ij = 0; check <dP,P> == <dZ,Z> (both equal sum G_ij s_ij)
                lhs, rhs = float((pred.grad * pred).sum()), float((z.grad * z).sum())
                euler_ok = abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
                flops = (2.0 * (b * k) ** 2 * e if all_steps else 2.0 * k * b * b * e) * 4        # fwd + recompute + dP + dZ
                print(json.dumps({"mode": "all-steps" if all_steps else "per-step", "N": n, "K": k, "B": b, "E": e,
                                  "fwd_bwd_ms": round(ms, 4), "tflops": round(flops / (ms * 1e-3) / 1e12, 2),
                                  "loss": round(float(loss), 5), "uniform_ok": uniform_ok, "euler_ok": euler_ok}))
                sys.stdout.flush()


if __name__ == "__main__":
    main()
