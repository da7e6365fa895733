#!/bin/bash
# One gpurun call worth of profiling for profiles/: plain runs first (each must exit 0 without ncu),
# then the ncu launch list of the bench command and ONE --set full capture of the top kernel.
# usage: scripts/profile_round.sh <tag>     (outputs under gpurun_out/)
set -x
TAG=${1:-r1}
OUT=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/bench_${TAG}_plain.json 2> $OUT/bench_${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/launches_${TAG}_vcycle.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_${TAG}_bench.log 2>&1
python scripts/probe_apply.py 1e8 4 > $OUT/plain_${TAG}_p4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_apply_affine --launch-skip 2 -c 1 \
    -o $OUT/prof_${TAG}_apply_p4 -f python scripts/probe_apply.py 1e8 4 > $OUT/ncu_${TAG}_p4.log 2>&1
tail -2 $OUT/ncu_${TAG}_p4.log
