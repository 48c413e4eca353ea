"""Text summaries of ncu outputs for profiles/ (run in the build container on files brought back in gpurun_out/).
usage: ncu_summary.py launches <launch-list.csv>        per-kernel totals of a `--metrics gpu__time_duration.sum` pass
       ncu_summary.py report <file.ncu-rep> [...]       key counters + stall picture of every kernel in `--set full` reports"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        name = r[k].split("(")[0].replace("void ", "").replace("tcc::", "")
        tot[name] += float(r[v].replace(",", "")) * 1e-6
        cnt[name] += 1
    s = sum(tot.values())
    print(f"{'kernel':55s} {'launches':>9s} {'total ms':>11s} {'share':>7s}")
    for name, t in tot.most_common():
        print(f"{name:55s} {cnt[name]:9d} {t:11.3f} {100 * t / s:6.1f}%")
    print(f"total {s:.1f} ms over {sum(cnt.values())} launches")


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {path}: {r[hdr.index('Kernel Name')][:110]}")
        for key in KEYS:
            if key in hdr:
                print(f"  {key:78s} {r[hdr.index(key)]} {units[hdr.index(key)]}")
        st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")]
        print("  warp stall reasons (warps per issue-active cycle): " +
              ", ".join(f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        for p in sys.argv[2:]:
            report(p)
