"""Probe (2 GPUs): which ways of putting an NCCL all-reduce inside a torch CUDA graph work with this torch / NCCL build.
usage: torchrun --nproc-per-node 2 tools/probes/nccl_capture_probe.py <variant>"""
import os
import sys
import traceback

import torch
import torch.distributed as dist

variant = sys.argv[1]
rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(1 << 20, device=dev) * (rank + 1)
w = torch.ones(1 << 20, device=dev, requires_grad=True)
comm = torch.cuda.Stream(device=dev)


def body():
    if variant in ("same_thread_global", "same_thread_local"):
        x.mul_(2.0)
        comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(comm):
            dist.all_reduce(x)
        torch.cuda.current_stream().wait_stream(comm)
        return x + 1
    if variant == "inline":
        x.mul_(2.0)
        dist.all_reduce(x)
        return x + 1
    if variant in ("hook_local", "hook_global"):
        w.grad = None
        holder = {}

        def hook(p):
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_reduce(p.grad)
        h = w.register_post_accumulate_grad_hook(hook)
        loss = (w * x).sum()
        loss.backward()
        h.remove()
        torch.cuda.current_stream().wait_stream(comm)
        return w.grad + 1
    raise SystemExit("unknown variant")


try:
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            out = body()
    torch.cuda.current_stream().wait_stream(side)
    dist.barrier()
    torch.cuda.synchronize()
    mode = "thread_local" if variant.endswith("local") else "global"
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode=mode):
        out = body()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print("VARIANT %s rank %d OK, out[0]=%.1f" % (variant, rank, float(out[0])), flush=True)
except Exception:
    print("VARIANT %s rank %d FAILED\n%s" % (variant, rank, traceback.format_exc()[-1800:]), flush=True)
    os._exit(1)
dist.destroy_process_group()
