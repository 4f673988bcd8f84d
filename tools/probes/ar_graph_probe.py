"""In-graph cost of the small-problem part of an e24 step: CUDA-graph replay time of forward + backward of the
autoregressive model alone (6 conv1d layers, 512/256 channels, T = 60 .. 1) at batch 64, and of the last two
encoder blocks alone.  Eager per-call event times overstate multi-launch calls; the launch list under ncu is cold-cache."""
import os
import sys
import time

import torch

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


def graph_time(fn, n=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


z = torch.randn(64, 512, 60, device=dev, requires_grad=True)
gz = torch.randn(64, 256, device=dev)


def ar_step():
    for p in model.autoregressive_model.parameters():
        p.grad = None
    z.grad = None
    c = model.autoregressive_model(z)
    c.backward(gz)


print("AR model fwd+bwd in a graph: %.3f ms" % graph_time(ar_step))

blocks = model.encoder.blocks
x2 = torch.randn(64, 128, 34, 156, device=dev, requires_grad=True)


def tail_step():
    for b in (blocks[2], blocks[3]):
        for p in b.parameters():
            p.grad = None
    x2.grad = None
    h = blocks[2](x2, outer_relu=True)
    h = blocks[3](h, outer_relu=False)
    h.sum().backward()


print("encoder blocks 2+3 fwd+bwd in a graph: %.3f ms" % graph_time(tail_step))
x3 = torch.randn(64, 256, 2, 77, device=dev, requires_grad=True)


def b3_step():
    for p in blocks[3].parameters():
        p.grad = None
    x3.grad = None
    blocks[3](x3, outer_relu=False).sum().backward()


print("encoder block 3 fwd+bwd in a graph: %.3f ms" % graph_time(b3_step))
os._exit(0)
