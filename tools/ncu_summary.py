"""Condense an .ncu-rep (ncu --set full) into the few metrics DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum", "local (spill) store bytes"),
    ("l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "local (spill) load bytes"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall long_scoreboard samples"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall barrier samples"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall short_scoreboard samples"),
    ("smsp__pcsamp_warps_issue_stalled_lg_throttle", "stall lg_throttle samples"),
    ("smsp__pcsamp_warps_issue_stalled_mio_throttle", "stall mio_throttle samples"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall math_pipe samples"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall wait samples"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall not_selected samples"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "issued (selected) samples"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path} (one block per captured launch; cold-cache, serialised)")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\n== {name}")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"  {label:32s} {r[i]} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
