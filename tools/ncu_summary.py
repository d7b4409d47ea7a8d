"""Summarise an .ncu-rep: one block per kernel launch with the metrics the roofline discussion uses,
plus (with --ops) the executed-instruction mix of the first matching kernel.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--ops] [--kernel regex]
"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy"),
    ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("smsp__inst_executed.sum", "warp_insts"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def ops(path, kernel):
    cmd = ["ncu", "-i", path, "--page", "source", "--csv"]
    if kernel:
        cmd += ["--kernel-name", "regex:" + kernel]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    hdr, cnt, samp, k = None, collections.Counter(), collections.Counter(), 0
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            k += 1
            if k == 1:
                print("instruction mix of:", r[1][:150])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if k != 1 or hdr is None or len(r) < 7:
            continue
        src = r[1].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        op = src.split()[0].split(".")[0] if src else "?"
        cnt[op] += int(r[hdr.index("Instructions Executed")] or 0)
        samp[op] += int(r[hdr.index("# Samples")] or 0)
    tot, ts = sum(cnt.values()), max(sum(samp.values()), 1)
    print(f"total warp instructions {tot}")
    for op, n in cnt.most_common(24):
        print(f"  {op:10s} {n:11d} {100 * n / tot:5.1f}%   stall samples {100 * samp[op] / ts:5.1f}%")


def main():
    path = sys.argv[1]
    kernel = None
    if "--kernel" in sys.argv:
        kernel = sys.argv[sys.argv.index("--kernel") + 1]
    hdr, units, rows = raw(path)
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        if kernel and kernel not in name:
            continue
        parts = []
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                parts.append(f"{short}={r[i]}{units[i] if short in ('time', 'dram_rd', 'dram_wr', 'smem_dyn') else ''}")
        print(" ".join(parts))
        print()
    if "--ops" in sys.argv:
        ops(path, kernel)


if __name__ == "__main__":
    main()
