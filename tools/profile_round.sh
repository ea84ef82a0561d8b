#!/bin/bash
# tools/profile_round.sh <tag> — runs ON THE GPU BOX (gpurun -- 'bash tools/profile_round.sh r01'): the bench lines of
# every workload, then the ncu launch list and one `--set full` capture of a 128-frame TUM1 step. Everything lands in
# gpurun_out/<tag>_*; tools/ncu_summary.py / tools/ncu_phases.py condense it into profiles/ afterwards.
# ncu runs only after the same command has exited 0 without it; numbers printed under ncu are never bench values.
T=${1:-r01}; O=gpurun_out; mkdir -p $O
for w in tum1 euroc kitti 4k kitti_stereo euroc_stereo euroc_rect tum1_frame tum1_track tum1_matchers; do
  extra=""; [ "$w" != tum1 ] && extra="--no-hamming"
  python bench.py --workload $w $extra > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err || echo "bench $w FAILED"
done
CMD="python bench.py --workload tum1 --batch 128 --steps 2 --warmup 3 --no-cpu-baseline --no-hamming"
if $CMD > $O/${T}_plain.log 2>&1; then
  ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 36 -c 24 --csv --log-file $O/${T}_launches_raw.csv $CMD > $O/${T}_ncu_launch.log 2>&1
  ncu --set full --clock-control none --import-source on --launch-skip 36 -c 12 -f -o $O/${T}_full $CMD > $O/${T}_ncu_full.log 2>&1
else
  echo "plain run failed"; tail -5 $O/${T}_plain.log
fi
ls -la $O | tail -20
