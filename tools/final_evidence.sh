#!/bin/bash
# Round-end evidence run (one GPU): bench lines, launch list with DRAM bytes, ncu --set full of the kernels DESIGN.md cites.
# Every program first runs WITHOUT ncu (its exit status gates the capture); numbers printed under ncu are never bench values.
set -u
O=gpurun_out
PF="python tools/profile_forward.py --precision bf16 --iters 1"
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02_default_bench.json 2> $O/r02_default_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_reference_arm.json 2> $O/r02_reference_arm.err; echo "ref rc=$?"
timeout 100 python tools/trace_forward.py > $O/r02_c3_bf16_trace.txt; echo "trace rc=$?"
timeout 120 $PF > $O/plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02_c3_bf16_launches.csv $PF > $O/ncu_ll.log 2>&1
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $O/r02_$1 $PF > $O/ncu_$1.log 2>&1; echo "$1 rc=$?"
}
cap convtc_refine_c3 k_conv_tc 49
cap convtc_32x32_c3 k_conv_tc 44
cap edge_chain_l0 k_edge_chain 3
cap lka_tail64 'k_lka_tail\(' 1
cap lka_tail128 k_lka_tail128 1
cap modulate_hr4 k_modulate_hr4 1
cap fft2_cols k_fft2_cols 1
timeout 120 python tools/profile_drct.py bf16 > $O/plain_drct.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_window_attn_tc -s 125 -c 1 -o $O/r02_window_attn_tc python tools/profile_drct.py bf16 > $O/ncu_wattn.log 2>&1; echo "wattn rc=$?"
ls -la $O/*.ncu-rep
