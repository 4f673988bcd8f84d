#!/usr/bin/env python
"""One fused InfoNCE forward + backward at a sweep size, for ncu:  python tools/profile_infonce.py [B K E all_steps]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))
import torch
import cpc_b200

b, k, e, all_steps = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (2048, 8, 512, 0)))
g = torch.Generator(device="cuda").manual_seed(0)
pred = (torch.randn(b, k, e, generator=g, device="cuda") / math.sqrt(e)).requires_grad_(True)
z = torch.randn(b, e, k, generator=g, device="cuda").requires_grad_(True)
for _ in range(2):
    pred.grad = z.grad = None
    loss = cpc_b200.ops.infonce(pred, z, bool(all_steps), "linear", 0.0)[0]
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
