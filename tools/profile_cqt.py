"""Runs the fused CQT front end alone (B = 64, L = 97 024, phase mode) a few times: target for `ncu --set full -k regex:cqt_umma`."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))
import cpc_b200                                              # noqa: E402
dev = torch.device("cuda:0")
pre = cpc_b200.PreprocessingModule(dict(cpc_b200.cqt_default_dict), phase=True).to(dev)
x = 0.1 * torch.randn(64, 1, 97024, device=dev)
for _ in range(3):
    y = pre(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = pre(x)
e1.record()
torch.cuda.synchronize()
print("cqt front end: %.3f ms per call" % (e0.elapsed_time(e1) / 10))
