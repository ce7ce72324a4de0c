"""Summarise an ncu report (read on the CPU box) into profiles/:
    python scripts/ncu_summary.py <rep> <out.txt> <users> [note] [items] [dim] [k]"""
import csv
import io
import subprocess
import sys

rep, out_path, users = sys.argv[1], sys.argv[2], int(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else ""
items = int(sys.argv[5]) if len(sys.argv) > 5 else 1000000
dim = int(sys.argv[6]) if len(sys.argv) > 6 else 128
k_arg = int(sys.argv[7]) if len(sys.argv) > 7 else 10
d_pad = -(-dim // 64) * 64
shard_mb = items * d_pad * 2 / 1e6  # 16-bit tensor-core copy
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
keys = [
    "Kernel Name", "Block Size", "Grid Size", "launch__registers_per_thread", "launch__cluster_dim_x", "gpu__time_duration.sum",
    "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_tmem_ldt.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
o = io.StringIO()
o.write("# ncu --set full --clock-control none --import-source on, one launch of the fused kernel\n")
if note:
    o.write("# " + note + "\n")
for k in keys:
    if k in d:
        o.write("%-88s %s %s\n" % (k, d[k], u.get(k, "")))
for k in hdr:
    if "stalled" in k and "per_issue_active" in k and float(d[k] or 0) > 0.2:
        o.write("%-88s %s\n" % (k, d[k]))
t_ms = float(d["gpu__time_duration.sum"]) * (1e-3 if u["gpu__time_duration.sum"].startswith("us") else 1.0)
flops = 2.0 * users * items * dim
def to_bytes(key):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[key]]
    return float(d[key]) * scale


dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
waves = -(-(users // 256 + (1 if users % 256 else 0)) // 74)
o.write(
    "\nderived: algorithmic FLOP = 2*U*N*d = %.3e -> %.0f TFLOP/s under ncu (clock-control none, cold caches); "
    "traffic = dram read + write = %.3f GB per launch = %.2f x (%.0f MB 16-bit item shard) for %d waves of subject tiles "
    "(74 CTA pairs stream the shard in lock-step: one HBM pass per wave, the other 73 reads are L2 hits); "
    "compulsory one-pass traffic would be %.0f MB\n"
    % (flops, flops / (t_ms * 1e-3) / 1e12, dram / 1e9, dram / (shard_mb * 1e6), shard_mb, waves, shard_mb + users * d_pad * 2 / 1e6)
)
import json

json.dump({"users": users, "items": items, "dim": dim, "k": k_arg, "dram_bytes_per_launch": dram, "kernel_ms_under_ncu": t_ms,
           "source": out_path}, open(out_path.replace(".txt", ".json"), "w"))
open(out_path, "w").write(o.getvalue())
print(o.getvalue())
