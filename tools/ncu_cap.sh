#!/bin/bash
# ncu --set full capture of named kernels inside one bf16 C3 forward: tools/ncu_cap.sh name:regex:skip ...   (outputs gpurun_out/r02_<name>.ncu-rep)
O=gpurun_out
PF="python tools/profile_forward.py --precision bf16 --iters 1"
timeout 120 $PF > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
for spec in "$@"; do
  IFS=: read name regex skip <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o $O/r02_$name $PF > $O/ncu_$name.log 2>&1; echo "$name rc=$?"
done
