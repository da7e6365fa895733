#!/bin/bash
# sweep TMA kernel tuning (threads per CTA x ring depth); usage: sweep_tma.sh "4 3 2 1"
for P in ${1:-4 3 2 1}; do
  for T in 128 64; do
    for R in 2 3 4; do
      if [ $T = 128 ] && [ $R = 4 ]; then continue; fi
      echo -n "P=$P tpb=$T R=$R  "
      PMGX_TMA_TPB=$T PMGX_TMA_R=$R timeout 120 python scripts/probe_apply.py 1e8 $P 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms'],3), round(d['gdofs'],2), round(d['frac'],3))" 2>&1 | tail -1
    done
  done
done
