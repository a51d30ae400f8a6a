"""Turns ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_banded_x2_full.md [traffic.json key]
  python tools/ncu_summary.py full gpurun_out/prof.raw.csv profiles/...md      (raw page exported on the box)
"""
import collections
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write(f"source: `{src}`; per-launch times are cold-cache and serialised: compare SHARES.\n\n")
        f.write("| kernel | launches | mean us | share of captured time |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:110]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |\n")
        # the kernels of ONE evaluation step: this library's kernels that launch as often as the
        # dominant row-streaming kernel (micro-benchmarks, plan set-up and torch kernels excluded)
        mine = {k: v for k, v in agg.items() if "lslb::" in k and "peak_kernel" not in k and "probe" not in k}
        if mine:
            top = max(mine.items(), key=lambda kv: sum(kv[1]))
            step = {k: v for k, v in mine.items() if len(v) == len(top[1])}
            st = sum(sum(v) / len(v) for v in step.values())
            f.write("\n## One evaluation step (kernels launched once per step)\n\n")
            f.write("| kernel | mean us | share of the step |\n|---|---:|---:|\n")
            for k, v in sorted(step.items(), key=lambda kv: -sum(kv[1])):
                f.write(f"| `{k[:90]}` | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / len(v) / st:.3f} |\n")
    print("wrote", dst)


def full(src, dst, traffic_key=None):
    # src: an .ncu-rep, or the CSV of its raw page already exported on the GPU box
    # (`ncu -i x.ncu-rep --page raw --csv > x.raw.csv`: reports with source can exceed what gpurun brings back)
    if src.endswith(".csv"):
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[H.index("Kernel Name")], "id": r[H.index("ID")]}
        for k in KEEP:
            if k in H:
                d[k] = (r[H.index(k)], U[H.index(k)])
        out.append(d)
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary\n\nsource: `{src}`\n\n")
        for d in out:
            f.write(f"## launch {d['id']}: `{d['kernel'][:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEEP:
                if k in d:
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            f.write("\n")
    if traffic_key and out:
        def tobytes(v, u):
            x = float(v.replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tr = sum(tobytes(*out[0][k]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        path = "profiles/traffic.json"
        try:
            cur = json.load(open(path))
        except (OSError, ValueError):
            cur = {}
        cur[traffic_key] = tr
        cur[traffic_key + "_source"] = dst
        json.dump(cur, open(path, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
