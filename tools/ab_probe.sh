#!/bin/bash
# development A/B: run tools/c5_probe.py for every variant library under incompressible-flow_b200/variants/
# usage: tools/ab_probe.sh <n> <tag>...   (tag "base" = the in-tree library)
n=$1; shift
for t in "$@"; do
  if [ "$t" = base ]; then unset IMPES_B200_LIB; else export IMPES_B200_LIB=$PWD/incompressible-flow_b200/variants/lib_$t.so; fi
  echo "=== $t"
  timeout 200 python tools/c5_probe.py $n 1 1 2>&1 | grep -E "step 1|breakdown|flush=1"
done
