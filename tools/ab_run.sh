#!/bin/bash
# Run bench.py once per experiment library and print symbols/s: tools/ab_run.sh [streams]   (on the GPU box)
cd "$(dirname "$0")/.."
for so in ofdm-course_b200/lib/exp/*.so; do
  v=$(OFDM_B200_LIB=$PWD/$so python bench.py --no-cpu --steps 10 --warmup 3 --streams ${1:-32768} --e2e-streams 512 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f M sym/s  ber %.3e' % (d['value']/1e6, d['config']['ber']))")
  echo "$(basename $so): $v"
done
