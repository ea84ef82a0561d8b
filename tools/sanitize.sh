#!/bin/bash
# tools/sanitize.sh <tag> — runs ON THE GPU BOX: compute-sanitizer memcheck over one pass of the hot path (the smoke test of
# __graft_entry__: extractor + Hamming + undistort + BoW + projection / local-points matchers) and racecheck over the
# extractor alone. Output lands in gpurun_out/<tag>_memcheck.log / _racecheck.log; the summary lines go to profiles/.
T=${1:-r02}; O=gpurun_out; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > $O/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/${T}_memcheck.log
cat > /tmp/race_one.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from orb_slam2_commit_b200 import ORBextractor, synth
img = synth.synth_image(320, 240, 99)
ex = ORBextractor(300, 1.2, 4, 20, 7, device=0)
k, d = ex(img)
print("keypoints", len(k))
PY
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/race_one.py > $O/${T}_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/${T}_racecheck.log
tail -4 $O/${T}_memcheck.log; tail -6 $O/${T}_racecheck.log
