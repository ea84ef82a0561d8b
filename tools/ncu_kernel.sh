#!/bin/bash
# tools/ncu_kernel.sh <tag> <kernel-regex> [batch] — runs ON THE GPU BOX: one `ncu --set full` capture of the named kernel(s)
# inside a resident TUM1 step (tools/exp_step.py), after the same command has exited 0 without ncu.
T=$1; K=$2; B=${3:-128}; O=gpurun_out; mkdir -p $O
CMD="python tools/exp_step.py $B 2 tum1"
if $CMD > $O/${T}_plain.log 2>&1; then
  ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip ${SKIP:-3} -c ${COUNT:-1} -f -o $O/${T} $CMD > $O/${T}_ncu.log 2>&1
  tail -2 $O/${T}_ncu.log
else
  echo "plain run failed"; tail -5 $O/${T}_plain.log
fi
