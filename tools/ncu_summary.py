"""Summarises gpurun_out/<tag>_launches.csv and <tag>_<kernel>.ncu-rep into profiles/<tag>_*.txt
(the tracked evidence; gpurun_out/ is scratch).  Usage: python tools/ncu_summary.py r1a knn2"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, kern = sys.argv[1], sys.argv[2]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

launches = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
    tot, cnt, shape = collections.Counter(), collections.Counter(), {}
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        tot[name] += v
        cnt[name] += 1
        shape[name] = (r[gi], r[bi])
    s = sum(tot.values())
    with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): "
                f"per-kernel totals of {launches.replace(ROOT + '/', '')}\n")
        f.write(f"{'kernel':44s} {'launches':>8s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}  grid/block\n")
        for k, v in tot.most_common():
            f.write(f"{k:44s} {cnt[k]:8d} {v / 1e3:12.1f} {v / 1e3 / cnt[k]:10.1f} {v / s:7.3f}  {shape[k][0]} {shape[k][1]}\n")
    print(open(os.path.join(out_dir, f"{tag}_launches.txt")).read())

rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{kern}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard", "smsp__warps_issue_stalled"]
    with open(os.path.join(out_dir, f"{tag}_{kern}_metrics.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, raw page of {rep.replace(ROOT + '/', '')} "
                f"({len(rows) - 2} captured launches)\n")
        for i, h in enumerate(hdr):
            if any(h == w or (w.endswith("stalled") and w in h) for w in want):
                f.write(f"{h} [{units[i]}]: {', '.join(r[i] for r in rows[2:])}\n")
    print(open(os.path.join(out_dir, f"{tag}_{kern}_metrics.txt")).read())
    # top stall lines of the source page (needs -lineinfo)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    with open(os.path.join(out_dir, f"{tag}_{kern}_sass_hot.txt"), "w") as f:
        lines = list(csv.reader(src.splitlines()))
        h = None
        body = []
        for r in lines:
            if "Source" in r and "# Samples" in " ".join(r) or ("Source" in r and "Warp Stall Sampling (All Samples)" in r):
                h = r
                continue
            if h and len(r) == len(h):
                body.append(r)
        if h:
            si = h.index("Source")
            ci = h.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in h else None
            if ci is not None:
                tot_s = sum(float(r[ci] or 0) for r in body) or 1.0
                top = sorted(body, key=lambda r: -float(r[ci] or 0))[:40]
                f.write("# hottest SASS instructions by warp-stall samples (first captured launch)\n")
                for r in top:
                    f.write(f"{float(r[ci] or 0) / tot_s:7.3%}  {r[si]}\n")
    print(open(os.path.join(out_dir, f"{tag}_{kern}_sass_hot.txt")).read()[:3000])
