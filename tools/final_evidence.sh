#!/bin/bash
# Round-end evidence run (one GPU): tests, bench lines, launch list with DRAM bytes, ncu --set full of the kernels DESIGN.md cites.
# Every program first runs WITHOUT ncu (its exit status gates the capture); numbers printed under ncu are never bench values.
set -u
O=gpurun_out
PF="python tools/profile_forward.py --precision bf16 --iters 1"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -q -m gpu > $O/r02_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_gpu_tests.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02_default_bench.json 2> $O/r02_default_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_reference_arm.json 2> $O/r02_reference_arm.err; echo "ref rc=$?"
timeout 100 python tools/trace_forward.py > $O/r02_c3_bf16_trace.txt; echo "trace rc=$?"; tail -1 $O/r02_c3_bf16_trace.txt
timeout 120 $PF > $O/plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02_c3_bf16_launches.csv $PF > $O/ncu_ll.log 2>&1
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/r02_$1 $PF > $O/ncu_$1.log 2>&1; echo "$1 rc=$?"
  # gpurun brings back at most 64 MiB: keep the text pages, drop the report (except KEEP_REP ones)
  ncu -i $O/r02_$1.ncu-rep --page details > $O/r02_$1_ncu_full.txt 2>/dev/null
  python tools/ncu_raw.py $O/r02_$1.ncu-rep gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum smsp__inst_executed.sum issue_stalled sm__cycles_active.avg pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct 2>/dev/null | grep -v pcsamp >> $O/r02_$1_ncu_full.txt
  case " ${KEEP_REP:-convtc_refine_c3} " in *" $1 "*) ;; *) rm -f $O/r02_$1.ncu-rep;; esac
}
timeout 600 python bench.py --workload n1 --steps 3 --warmup 3 > $O/r02_n1_final.json 2> $O/r02_n1_final.err; echo "n1 rc=$?"
timeout 600 python bench.py --workload c5 --steps 2 --warmup 3 > $O/r02_c5_final.json 2> $O/r02_c5_final.err; echo "c5 rc=$?"
if [ "${SKIP_NCU:-0}" = "1" ]; then du -sh $O; exit 0; fi
cap convtc_refine_c3 k_conv_tc 39
cap convtc_32x32_c3 k_conv_tc 34
cap token_attn k_token_attn 1
cap token_ffn k_token_ffn 1
cap align_tokens k_align_tokens 1
cap crossband_attn2 k_crossband_attn2 1
cap selector k_selector 1
cap lka_tail64w k_lka_tailw 1
cap lka_tail128w k_lka_tail128w 1
cap lka_dw21_roll k_lka_dw21_roll 2
cap upsample_int k_upsample_int 5
cap spatial_gate_rows k_spatial_gate_rows 5
cap edge_chain_l0 k_edge_chain 3
cap modulate_hr4 k_modulate_hr4 1
du -sh $O
