#!/bin/bash
P=/root/repo/ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200
for v in "$@"; do
  echo "== $v"
  AIQMC_LIB=$P/dbg_$v.so python bench.py --walkers 32768 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('w-s/s %.0f  ms/step %.2f  quad share %.3f  quad ms %.2f' % (d['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['ms_per_step']*d['roofline']['share_of_step']))"
done
