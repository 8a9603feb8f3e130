#!/bin/bash
# Mid-round checkpoint on one GPU: the whole -m gpu suite, the default bench line, the in-situ launch trace.
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02_default_bench.json 2> $O/r02_default_bench.err; echo "bench rc=$?"
timeout 100 python tools/trace_forward.py > $O/r02_c3_bf16_trace.txt; echo "trace rc=$?"; tail -1 $O/r02_c3_bf16_trace.txt
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_default_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/img',d.get('ms_per_image_per_gpu'),'e2e',d['e2e']['value'],'parity',d.get('parity',{}).get('bf16'),'clocks',d.get('clocks'))
print('train',d.get('train',{}).get('value'), d.get('train',{}).get('ms_per_step'))
P
