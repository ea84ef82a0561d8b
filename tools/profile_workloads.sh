#!/bin/bash
# tools/profile_workloads.sh <tag> — runs ON THE GPU BOX: one `ncu --set full` capture of a resident step of the KITTI, EuRoC and
# 4K geometries (after the same command has exited 0 without ncu), so that bench.py's roofline.traffic / issue objects for
# those workloads come from their own capture instead of the TUM1 one. tools/ncu_summary.py full condenses them into
# gpurun_out/<tag>_ncu_full_<cfg>.csv on the box (the .ncu-rep files are deleted there: too large to bring back) (frames per launch: kitti 128, euroc 128, 4k 16).
T=${1:-r02}; O=gpurun_out; mkdir -p $O
for spec in "kitti 128 12" "euroc 128 12" "4k 16 16"; do
  set -- $spec; cfg=$1; B=$2; nk=$3
  CMD="python tools/exp_step.py $B 2 $cfg"
  if $CMD > $O/${T}_${cfg}_plain.log 2>&1; then
    ncu --set full --clock-control none --launch-skip $((3 * nk)) -c $nk -f -o $O/${T}_full_${cfg} $CMD > $O/${T}_${cfg}_ncu.log 2>&1
    python tools/ncu_summary.py full $O/${T}_full_${cfg}.ncu-rep > $O/${T}_ncu_full_${cfg}.csv 2>> $O/${T}_${cfg}_ncu.log
    rm -f $O/${T}_full_${cfg}.ncu-rep                     # the reports of these geometries exceed what gpurun brings back
    tail -1 $O/${T}_${cfg}_ncu.log; wc -l $O/${T}_ncu_full_${cfg}.csv
  else echo "plain run of $cfg failed"; tail -3 $O/${T}_${cfg}_plain.log; fi
done
