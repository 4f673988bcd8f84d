#!/usr/bin/env python
"""bench.py -- train audio-sec/sec of the CPC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE configs[1] = experiments['e24'] of the reference: CQT(+phase)
PreprocessingModule -> ScalogramResidualEncoder arch 7 -> ConvolutionalArModel arch 3 -> InfoNCE (linear,
all-steps, K=16), batch 64 items of 97 024 samples (6.064 s at 16 kHz) PER GPU (weak scaling), Adam step
included, synthetic white-noise audio, random-init weights.  One "step" = one full training step.

Our arm prints `value` (inputs resident in HBM), `e2e` (same step through the public trainer API with the
batch coming from pinned host memory and the loss read back every step), `roofline` of the dominant kernel
(timed with CUDA events inside real steps) and `cpu_baseline` (the oracle port of the same step on the host
cores, bounded sample).  `--impl reference` times that CPU port alone (the reference is CPU-only Python and
/root/reference does not travel to the GPU box, so kind = "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))

SR = 16000
BATCH_PER_GPU = 64
CPU_SAMPLE_BATCH = 8
METRIC = "train audio-sec/sec"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def count_since(self, t0):
        return sum(1 for t, _ in self.rows if t >= t0)

    def stop(self, t0=0.0):
        """Summary of the samples taken at or after wall-clock time t0 (the start of the measured load)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, r in self.rows:
            if t < t0:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_arm(steps, warmup, batch):
    """The oracle port of the full training step on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import cpc_oracle_model as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = M.OracleE24()
    g = torch.Generator().manual_seed(1234)
    x = 0.1 * torch.randn(batch, model.item_length, generator=g)
    M.train_steps(model, [x] * warmup)
    t0 = time.perf_counter()
    M.train_steps(model, [x] * steps)
    dt = (time.perf_counter() - t0) / steps
    audio_s = batch * model.item_length / SR
    return {"value": audio_s / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "sample": "%d steps of the same e24 training step at batch %d (%.1f audio-s per step), torch CPU fp32, "
                      "%d threads, anomaly mode off" % (steps, batch, audio_s, cores),
            "ms_per_step": dt * 1e3}


def workload_config(n_gpus):
    return {"workload": "e24: CQT(256 bins, phase) + ScalogramResidualEncoder arch7 + ConvolutionalArModel arch3 + "
                        "InfoNCE linear/all-steps K=16, full train step incl. Adam (BASELINE configs[1])",
            "batch_per_gpu": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * n_gpus, "samples_per_item": 97024,
            "sample_rate": SR, "parallelism": "dp%d" % n_gpus, "negatives": "per-GPU",
            "submission": "whole step captured in one CUDA graph (cpc_b200.GraphedTrainStep)",
            "l2": "activations (>300 MB per layer) exceed the 126 MB L2; no explicit flush"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    r = cpu_reference_arm(steps, warm, CPU_SAMPLE_BATCH)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.gpus), submission="CPU port of the reference training step (oracle/), "
                                                                  "batch %d per step" % CPU_SAMPLE_BATCH),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args):
    # stdout carries exactly one JSON line: NCCL's own banner (NCCL_DEBUG=VERSION) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    import cpc_b200
    from cpc_b200 import _lib, configs, ddp, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local = ddp.init_from_env("nccl")
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d; launch with torchrun for N>1" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.load().cpc_runtime_check(), "cpc_runtime_check")

    exp = configs.experiment("e24")
    tc = exp["training_config"]
    torch.manual_seed(0)
    model, pre, _ = configs.setup_model(exp["cqt_config"], exp["encoder_config"], exp["ar_model_config"], tc,
                                        device=dev)
    ddp.broadcast_parameters(model, 0)
    length = model.item_length
    b = BATCH_PER_GPU
    trainer = cpc_b200.ContrastiveEstimationTrainer(
        model=model, dataset=None, device=dev, regularization=tc["regularization"],
        score_over_all_timesteps=tc["score_over_all_timesteps"], score_function=tc["score_function"],
        preprocessing=pre, prediction_steps=tc["prediction_steps"], verbose=False)
    trainer.optimizer = tc["optimizer"]                        # torch.optim.Adam in the reference's config
    optimizer = trainer.make_optimizer(tc["learning_rate"])     # -> cpc_b200.optim.Adam (one kernel per step)
    use_graph = not args.no_graph
    reducer = ddp.GradientBucketReducer(model) if (world > 1 and not use_graph) else None
    model.train()

    g = torch.Generator().manual_seed(1234 + rank)
    host_batches = [(0.1 * torch.randn(b, length, generator=g)).pin_memory() for _ in range(2)]
    dev_batches = [h.to(dev) for h in host_batches]

    def eager_step(batch):
        loss, max_score = trainer.loss_on_batch(batch)
        model.zero_grad(set_to_none=True)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        elif world > 1:
            ddp.allreduce_gradients([p for p in model.parameters() if p.requires_grad], world)
        optimizer.step()
        return loss, max_score

    graphed, launches_per_step = None, None
    if use_graph:
        # the whole step (forward, backward, Adam) is captured once and replayed; count our launches per step while
        # capturing (a replay issues no host-side launch calls)
        _lib.reset_launch_count()
        graphed = cpc_b200.GraphedTrainStep(trainer, optimizer, (b, length), warmup=3)
        launches_per_step = _lib.launch_count() // 4            # 3 eager warm-up steps + the captured one

    def step(batch):
        return graphed(batch) if graphed is not None else eager_step(batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / k

    clocks = ClockSampler(local)
    clocks.start()                                              # nvidia-smi needs a moment to deliver its first sample
    for i in range(args.warmup):
        step(dev_batches[i % 2])
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    t_load = time.time()
    ms_dev = timed(lambda i: step(dev_batches[i % 2]), args.steps)
    launches = launches_per_step * args.steps if graphed is not None else _lib.launch_count()
    # K steps of ~16 ms are shorter than a few 100 ms sampling periods: keep the identical load running (untimed) until
    # ~0.6 s of it have been sampled.  The count derives from ms_dev (max over ranks), so every rank runs the same steps.
    clock_window = "timed region"
    n_extra = int(600.0 / max(ms_dev, 1e-3)) - args.steps
    if n_extra > 0:
        for i in range(n_extra):
            step(dev_batches[i % 2])
        torch.cuda.synchronize()
        clock_window = "timed region + %d identical untimed steps right after it" % n_extra
    clk = clocks.stop(t_load)
    clk["window"] = clock_window

    # end to end through the public API: pinned host batch -> device, full step, loss + max score read back
    last = {}

    # Every step's batch travels pinned host -> device inside the timed region.  Like a DataLoader with
    # pin_memory + non_blocking copies, the copy of batch i+1 is issued on a side stream while step i computes.
    copy_stream = torch.cuda.Stream(device=dev)
    staged = [torch.empty(b, length, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])             # the step that last read this staging buffer is done
            staged[i % 2].copy_(host_batches[i % 2], non_blocking=True)
            ready[i % 2].record(copy_stream)

    # The result of every step (loss, max score) is copied to pinned host memory and read by the host inside the timed
    # region -- one step late, like a training loop that logs asynchronously: the host reads step i-1 after it has
    # submitted step i, so the GPU is not idle while Python turns around (the last step is read before the clock stops).
    host_result = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    result_ready = [torch.cuda.Event() for _ in range(2)]

    def read_result(i):
        result_ready[i % 2].synchronize()
        last["v"] = host_result[i % 2].tolist()

    def e2e_step(i, n_steps=args.steps):
        if i == 0:
            prefetch(0)
        torch.cuda.current_stream(dev).wait_event(ready[i % 2])
        prefetch(i + 1)
        loss, mx = step(staged[i % 2])
        consumed[i % 2].record(torch.cuda.current_stream(dev))
        host_result[i % 2].copy_(torch.stack([loss.detach(), mx.detach()]), non_blocking=True)   # device -> host
        result_ready[i % 2].record(torch.cuda.current_stream(dev))
        if i > 0:
            read_result(i - 1)
        if i == n_steps - 1:
            read_result(i)
    for ev in consumed:
        ev.record(torch.cuda.current_stream(dev))
    e2e_step(0, 1)
    torch.cuda.synchronize()
    for ev in consumed:
        ev.record(torch.cuda.current_stream(dev))
    ms_e2e = timed(e2e_step, args.steps)

    # dominant kernel, timed with CUDA events around each C-ABI call inside real steps
    prof = ops.KernelProfiler()
    with prof:
        for i in range(max(2, min(args.steps, 4))):
            eager_step(dev_batches[i % 2])
    torch.cuda.synchronize()
    top = prof.summary()

    audio_s_step = world * b * length / SR
    value = audio_s_step / (ms_dev * 1e-3)
    e2e_value = audio_s_step / (ms_e2e * 1e-3)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = measured_peaks()
    # dominant kernel FAMILY (one CUDA kernel function serves several layer shapes): the family with the largest
    # share of the event-timed kernel time; achieved = sum of algorithmic work / sum of launch time
    fams = {}
    for k in top:
        name = k["key"].split("[")[-1].rstrip("]") if "[" in k["key"] else k["key"].split(" ")[0]
        f = fams.setdefault(name, {"family": name, "ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        f["ms"] += k["total_ms"]
        f["flops"] += k["flops_per_launch"] * k["count"]
        f["bytes"] += k["bytes_per_launch"] * k["count"]
        f["launches"] += k["count"]
    total_ms = sum(f["ms"] for f in fams.values()) or 1.0
    fam_list = sorted(fams.values(), key=lambda f: -f["ms"])
    for f in fam_list:
        f["share"] = f["ms"] / total_ms
        f["tflops"] = f["flops"] / (f["ms"] * 1e-3) / 1e12 if f["ms"] > 0 else 0.0
        f["gbs"] = f["bytes"] / (f["ms"] * 1e-3) / 1e9 if f["ms"] > 0 else 0.0
    roofline = None
    if fam_list:
        dom = fam_list[0]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(dom["family"], {}).get("dram_bytes_per_launch")
        if dom["flops"] > 0:
            roofline = {"bound": "tensor", "achieved": dom["tflops"], "peak": peaks["tensor"], "unit": "TFLOP/s",
                        "frac": dom["tflops"] / peaks["tensor"], "traffic": traffic, "kernel": dom["family"],
                        "avg_launch_ms": dom["ms"] / dom["launches"], "launches": dom["launches"],
                        "share_of_kernel_time": dom["share"],
                        "peak_source": peaks["source"] + " bf16 dense cuBLAS, sustained",
                        "note": "achieved = algorithmic conv FLOPs (2*Cout*Cin*kh*kw*B*OH*OW, summed over the family's "
                                "launches) / CUDA-event time; fp32-faithful mode issues 3 bf16 MMAs per product, so the "
                                "ceiling of `frac` is 1/3"}
        else:
            roofline = {"bound": "hbm", "achieved": dom["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": dom["gbs"] / peaks["hbm"], "traffic": traffic, "kernel": dom["family"],
                        "avg_launch_ms": dom["ms"] / dom["launches"], "launches": dom["launches"],
                        "share_of_kernel_time": dom["share"], "peak_source": peaks["source"] + " copy bandwidth",
                        "note": "achieved = algorithmic bytes (passes over the activation) / CUDA-event time"}
    cpu = cpu_reference_arm(2, 1, CPU_SAMPLE_BATCH)
    line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": b * length * 4, "d2h_bytes_per_step": 8, "last_loss": last["v"][0],
                    "result_read": "every step's (loss, max score) copied to pinned host memory and read by the host "
                                   "inside the timed region, one step late (asynchronous logging)"},
            "roofline": roofline,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "kernel_families": [{k: f[k] for k in ("family", "share", "ms", "launches", "tflops", "gbs")} for f in fam_list],
            "kernels": top[:40]}
    emit(line)


_RESULT_FD = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries loaded later write banners to file descriptor 1 (NCCL
    prints "NCCL version ..." there at communicator creation).  Keep a private duplicate of the real stdout for the
    result line and point descriptor 1 (and sys.stdout) at stderr for everything else."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="submit kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
