#!/bin/bash
# A/B of environment switches on the in-situ launch trace: tools/ab_trace.sh "grep-regex" "ENV=val ..." "ENV2=val" ...   ("-" = no switch)
re=$1; shift
for envs in "$@"; do
  echo "== $envs"
  if [ "$envs" = "-" ]; then envs=""; fi
  env $envs timeout 100 python tools/trace_forward.py 2>&1 | grep -E "$re|sum of"
done
