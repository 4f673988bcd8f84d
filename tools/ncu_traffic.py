"""profiles/ncu_traffic.json from an `ncu --set full` report of tools/profile_kernel.py:

    python tools/ncu_traffic.py <report.ncu-rep> "<bench kernel key>" <kernel-name-substring> [summary.md]

Takes dram__bytes_read.sum + dram__bytes_write.sum of the LAST launch whose name contains the substring, stamps the
entry with the digest of the kernel sources (build._digest()) so that bench.py refuses it after the kernels change, and
appends a short metric summary to the markdown file."""
import csv
import io
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "constrastive-predictive-coding-audio_b200"))
import build                                                 # noqa: E402

report, key, needle = sys.argv[1:4]
summary = sys.argv[4] if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
picked = [r for r in rows[2:] if needle in r[col["Kernel Name"]]]
if not picked:
    raise SystemExit("no launch matching %r in %s" % (needle, report))
r = picked[-1]


def val(name):
    v = r[col[name]].replace(",", "")
    u = units[col[name]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9,
             "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "second": 1.0}.get(u, 1.0)
    return float(v) * scale


dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
data = {"kernels": {}}
digest = build._digest()
if os.path.exists(path):
    with open(path) as fh:
        old = json.load(fh)
    if old.get("source_digest") == digest:
        data = old
data["source_digest"] = digest
data["captured"] = time.strftime("%Y-%m-%d") + " round 2"
data["kernels"][key] = {"dram_bytes_per_launch": dram, "kernel": r[col["Kernel Name"]][:80],
                        "duration_s_under_ncu": val("gpu__time_duration.sum")}
with open(path, "w") as fh:
    json.dump(data, fh, indent=1)
print(key, "->", dram, "bytes per launch")
if summary:
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    with open(summary, "a") as fh:
        fh.write("\n## %s\n\nkernel `%s` (last matching launch of `%s`)\n\n| metric | value | unit |\n|---|---:|---|\n"
                 % (key, r[col["Kernel Name"]][:90], os.path.basename(report)))
        for name in want:
            hits = [h for h in hdr if h.endswith(name) or h == name]
            for h in hits[:1]:
                fh.write("| %s | %s | %s |\n" % (h, r[col[h]], units[col[h]]))
