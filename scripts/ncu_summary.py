"""Prints the headline counters of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg.per_second"]
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True,
                                      text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
extra = sys.argv[2:]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for h, u in zip(hdr, units):
        if h in WANT or any(e in h for e in extra):
            if d[h] not in ("", "0"):
                print(f"   {h} [{u}] = {d[h]}")
