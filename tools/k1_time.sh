#!/bin/bash
# LAB: K1 + K2 device times of the c2 bench line, twice (same box), no CPU leg / extras
for i in 1 2; do
python bench.py --no-extras --no-cpu --c3 multi 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print('cells/s %.4g  step %.4f ms  K1 %.4f ms  K2 %.4f ms  frac %.3f' % (l['value'], l['ms_per_step'], r['ms'], r['eig_ms'], r['frac']))"
done
