#!/bin/bash
# Final measurements of a round for profiles/: driver-style bench at N GPUs (+ reference arm at N=1),
# C++ drivers.  usage: scripts/final_round.sh <N> <tag>
N=${1:-1}; TAG=${2:-r1}; OUT=gpurun_out
if [ "$N" = 1 ]; then
  python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $OUT/bench_${TAG}_ref_n1.json 2> $OUT/bench_${TAG}_ref_n1.err
  python bench.py --gpus 1 --steps 5 --warmup 3 > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
  ./examples/build/pmg_main --ndofs 1e8 --niter 10 > $OUT/cpp_pmg_${TAG}_n1.log 2>&1
  ./examples/build/cg_main --ndofs 2e8 --degree 6 > $OUT/cpp_cg_${TAG}_n1.log 2>&1
  tail -4 $OUT/cpp_pmg_${TAG}_n1.log; tail -3 $OUT/cpp_cg_${TAG}_n1.log
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N \
      bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
  rm -f /tmp/pmgx_id_$N
  for r in $(seq 0 $((N-1))); do
    RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r ./examples/build/pmg_main --ndofs 2e7 --niter 4 --idfile /tmp/pmgx_id_$N > $OUT/cpp_pmg_${TAG}_n${N}_r$r.log 2>&1 &
  done
  wait
  tail -3 $OUT/cpp_pmg_${TAG}_n${N}_r0.log
fi
python - <<PY
import json
d=json.loads(open("$OUT/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.3f Gdof/s  %.2f ms/cycle  apply %.3f ms (%.1f%%)  roofline %.3f  e2e %.3f Gdof/s (%.2f ms)  coarse its %s  halo %s" % (
  d["n_gpus"], d["value"], d["ms_per_step"], d["apply"]["ms"], 100*d["apply"]["frac_of_hbm_peak"], d["roofline"]["frac"],
  d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["coarse_iterations_last_cycle"], d["config"]["halo"]))
PY
