"""Print selected metrics from `ncu -i rep --page raw --csv` (reads the csv on stdin or a path)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
        "sm__warps_active.avg.pct", "launch__registers_per_thread", "sm__throughput.avg.pct",
        "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct", "launch__grid_size", "sm__cycles_elapsed.avg",
        "l1tex__throughput", "smsp__cycles_active.avg", "sm__cycles_active.avg", "dram__throughput",
        "lts__t_sector_hit_rate", "sm__pipe_tensor_subpipe", "smsp__inst_executed.sum", "launch__occupancy_limit",
        "smsp__warp_issue_stalled", "l1tex__m_xbar2l1tex_read_bytes", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_bytes_equiv_l1sectormiss", "sm__sass_inst_executed_op_shared", "smsp__pcsamp"]
for i, h in enumerate(hdr):
    if any(h.startswith(k) for k in keys):
        print(f"{h:80s}", [r[i] for r in rows[1:] if len(r) > i])
