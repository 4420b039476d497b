#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) -- or the `ncu -i rep --page raw --csv` dump made on the box -- into the counters
that matter for the step and learner kernels.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt
    python tools/ncu_summary.py gpurun_out/prof.raw.csv > profiles/rNN_name.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active"]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {rep}: {len(data)} launch(es) of", data[0][hdr.index('Kernel Name')][:90])
    keys = KEYS + [h for h in hdr if "tensor" in h and h not in KEYS and "ops_path" not in h and "attribute" not in h
                   and ".avg" in h and ("pct_of_peak_sustained_active" in h or h.endswith(".avg"))]
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {units[i]:12s} " + "  ".join(r[i] for r in data))
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls.append((float(data[0][i]), h))
            except ValueError:
                pass
    print("# warp stall reasons (warps per issue-active cycle), first launch")
    for v, h in sorted(stalls, reverse=True)[:10]:
        print(f"{v:8.3f}  {h}")


if __name__ == "__main__":
    main()
