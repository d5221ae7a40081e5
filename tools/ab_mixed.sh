#!/bin/bash
# A/B of prebuilt library variants and run-time options on one GPU box:
#   tools/ab_mixed.sh "variants/a.so" "variants/b.so OFDG_SHADE_BLOCKS_PER_SM=5" "- OFDG_PIPELINE=0" ...
# (first word: library to copy over csrc/libofdg.so, "-" = the one in place; rest: environment for bench.py)
LIB=optical-flow-2d-data-generation_b200/csrc/libofdg.so
cp $LIB /tmp/libofdg_current.so
for spec in "$@"; do
  set -- $spec
  v=$1; shift
  if [ "$v" = "-" ]; then cp /tmp/libofdg_current.so $LIB; else cp "$v" $LIB; fi
  for rep in $(seq 1 ${REPS:-1}); do
    env "$@" python bench.py --no-cpu --no-layer --no-other-configs --e2e-steps 3 --steps ${STEPS:-400} 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']; k = r.get('kernels') or {}
print('[$spec]', 'samples/s %.0f' % d['value'], 'step %.4f ms' % d['ms_per_step'], 'prod %.0f' % d['production_mode']['value'],
      ' '.join('%s %.4f' % (n.replace('_kernel', ''), v['ms']) for n, v in k.items()), 'serial %.4f' % (r.get('serial_step_ms') or 0))"
  done
done
cp /tmp/libofdg_current.so $LIB
