"""Runs ONE C-ABI kernel at the shape named by a bench.py kernel key, a few times -- the target of
`ncu --set full -k regex:<kernel>` for the roofline's `traffic` figure (tools/ncu_traffic.py turns the report into
profiles/ncu_traffic.json).  Keys look like
    cpc_conv_dgrad b64 128x63x156->128x34x156 k30x1 s1x1 [tall_conv_tcgen05_128ch]
    cpc_cqt_fwd b64 L97024 T630 mode2
"""
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))
import cpc_b200                                              # noqa: E402
from cpc_b200 import ops                                     # noqa: E402

key = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
m = re.match(r"cpc_conv_(fwd|dgrad|wgrad) b(\d+) (\d+)x(\d+)x(\d+)->(\d+)x(\d+)x(\d+) k(\d+)x(\d+) s(\d+)x(\d+)", key)
if m:
    which = m.group(1)
    b, ci, h, w, co, oh, ow, kh, kw, sh, sw = (int(v) for v in m.groups()[1:])
    top = (oh - 1) * sh + kh - h                                # zero rows above the input (ZeroPad2d top padding)
    x = torch.randn(b, ci, h, w, device=dev, requires_grad=which != "fwd")
    wt = (torch.randn(co, ci, kh, kw, device=dev) / (ci * kh * kw) ** 0.5).requires_grad_(which != "fwd")
    gy = torch.randn(b, co, oh, ow, device=dev)
    with ops.KernelProfiler() as prof:
        for _ in range(reps):
            y = ops.conv2d(x, wt, None, (sh, sw), (0, 0), extra_top=max(top, 0))
            assert tuple(y.shape) == (b, co, oh, ow), (tuple(y.shape), (b, co, oh, ow))
            if which != "fwd":
                y.backward(gy)
                x.grad = wt.grad = None
    for k in prof.summary():
        print("%-100s avg %.3f ms  %.1f TFLOP/s" % (k["key"], k["avg_ms"], k["tflops"]))
elif key.startswith("cpc_cqt_fwd"):
    m = re.match(r"cpc_cqt_fwd b(\d+) L(\d+) T(\d+) mode(\d)", key)
    b, l, _, mode = (int(v) for v in m.groups())
    pre = cpc_b200.PreprocessingModule(dict(cpc_b200.cqt_default_dict), phase=mode == 2).to(dev)
    x = 0.1 * torch.randn(b, 1, l, device=dev)
    with ops.KernelProfiler() as prof:
        for _ in range(reps):
            pre(x)
    for k in prof.summary():
        print("%-100s avg %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (k["key"], k["avg_ms"], k["tflops"], k["gbs"]))
else:
    raise SystemExit("unsupported key: " + key)
torch.cuda.synchronize()
