"""Summarise an ncu --set full report (read here, no GPU needed) into profiles/<tag>_ncu_summary.md and refresh
profiles/render_kernel_traffic.json (the numbers bench.py quotes in `roofline`).   python scripts/ncu_summary.py <rep> <tag>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles")  # on the GPU box: gpurun_out (the only directory copied back)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
out = [f"# ncu --set full summary, {tag} (`scripts/profile.sh`, bench.py --steps 2 --warmup 1)\n",
       "Captured with `--clock-control none`; durations here are profiler-serialised single launches, not bench values.\n"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = name.split("(")[0]
    out.append(f"\n## {short}\n\n| metric | value | unit |\n|---|---|---|")
    vals = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {r[i]} | {units[i]} |")
            vals[w] = (r[i], units[i])
    if "render_kernel" in name and "bwd" not in name:
        def num(k):
            v, u = vals[k]
            v = float(v.replace(",", ""))
            return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)
        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        js = {"kernel": short, "ncu_source": f"profiles/{tag}_ncu_summary.md (ncu --set full, scripts/profile.sh)",
              "dram_bytes_read_per_launch": int(rd), "dram_bytes_write_per_launch": int(wr), "dram_bytes_per_launch": int(rd + wr),
              "l1_wavefront_pct": float(vals["l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"][0]),
              "warps_active_pct": float(vals["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
              "note": "65536 rays x 32 samples, L2 flushed before the step"}
        with open(os.path.join(OUT, "render_kernel_traffic.json" if tag.count("_") == 0 else f"{tag}_traffic.json"), "w") as f:
            json.dump(js, f, indent=1)
with open(os.path.join(OUT, f"{tag}_ncu_summary.md"), "w") as f:
    f.write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}_ncu_summary.md")
