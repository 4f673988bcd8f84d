"""Kernel table (torch.profiler / CUPTI) of forward + backward of the e24 autoregressive model alone at batch 64."""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))
import cpc_b200                                              # noqa: E402
from cpc_b200 import configs                                 # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
exp = configs.experiment("e24")
tc = exp["training_config"]
torch.manual_seed(0)
model, pre, _ = configs.setup_model(exp["cqt_config"], exp["encoder_config"], exp["ar_model_config"], tc, device=dev)
model.train()
z = torch.randn(64, 512, 60, device=dev, requires_grad=True)
gz = torch.randn(64, 256, device=dev)


def ar_step():
    for p in model.autoregressive_model.parameters():
        p.grad = None
    z.grad = None
    model.autoregressive_model(z).backward(gz)


for _ in range(3):
    ar_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ar_step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
n = sum(e.count for e in rows)
print("AR fwd+bwd: %d device activities, %.1f us of device time" % (n, total))
for e in rows:
    print("%-90s n=%-3d %8.1f us" % (e.key[:90], e.count, e.device_time_total))
os._exit(0)
