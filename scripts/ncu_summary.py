"""Summarise an `ncu --set full` report of scripts/ncu_step.py into profiles/r2/ (run in the build container):
    python scripts/ncu_summary.py gpurun_out/ncu_full_r2.ncu-rep
writes profiles/r2/ncu_full_kernels.txt (per-kernel table) and profiles/r2/traffic_bf16.json (DRAM bytes per launch of
the iteration kernel and the token stage, keyed by the digest of the kernel sources they were captured from)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory % (elapsed)"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem"), ("smsp__inst_executed.sum", "warp instructions")]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


out = ["ncu --set full --clock-control none, scripts/ncu_step.py (third eager SLATE encode step, batch 64, bf16 mode), B200",
       "one launch per row, in launch order; cold-cache, serialised (compare shares / traffic, not absolute times)", ""]
traffic = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "")
    out.append(short)
    for c, label in cols:
        if c in idx:
            out.append(f"    {label:28s} {r[idx[c]]} {units[idx[c]]}")
    rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    key = "sa_iter_fwd" if "sa_iter_fwd" in name else "kv_proj_fwd" if "kv_proj_tc" in name else None
    if key:
        traffic[key] = {"kernel": short, "traffic_bytes": rd + wr, "read": rd, "write": wr,
                        "time_us": float(r[idx["gpu__time_duration.sum"]])}
digest = open(os.path.join(ROOT, "ocrl_b200", "csrc", "build", "digest.txt")).read().strip()
traffic["build_digest"] = digest
traffic["config"] = [64, 4096, 192, 6, 3]
traffic["source"] = "profiles/r2/ncu_full_kernels.txt (ncu --set full, scripts/ncu_step.py)"
os.makedirs(os.path.join(ROOT, "profiles", "r2"), exist_ok=True)
open(os.path.join(ROOT, "profiles", "r2", "ncu_full_kernels.txt"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "r2", "traffic_bf16.json"), "w"), indent=1)
print("\n".join(out[:8]))
print(json.dumps(traffic)[:400])
