"""Text summary of an `ncu --set full` report for profiles/: selected metrics per captured launch.
usage: python tools/ncu_summary.py report.ncu-rep [label ...]   (labels name the launches in capture order)"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    labels = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    for n, d in enumerate(data):
        name = d[kn].replace("ssnb::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = name.split("(")[0] if not name.startswith("void") else name[5:].split("(CU")[0]
        print("== launch %d: %s%s" % (n, name, ("  [" + labels[n] + "]") if n < len(labels) else ""))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("   %-100s %s %s" % (m, d[i], units[i]))
        print()


if __name__ == "__main__":
    main()
