#!/bin/bash
# A/B of prebuilt library variants on one GPU box: tools/ab_variants.sh variants/libofdg_a.so variants/libofdg_b.so ...
# (each variant is copied over csrc/libofdg.so, bench.py runs without the CPU leg; one summary line per variant and run;
# the library that was in place comes last). REPS=1 for a single run each.
LIB=optical-flow-2d-data-generation_b200/csrc/libofdg.so
cp $LIB /tmp/libofdg_current.so
for v in "$@" /tmp/libofdg_current.so; do
  cp "$v" $LIB
  for rep in $(seq 1 ${REPS:-2}); do
    python bench.py --no-cpu --no-layer --no-other-configs --e2e-steps 3 --steps ${STEPS:-600} 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']; k = r.get('kernels') or {}
print('$v', 'samples/s %.0f' % d['value'], 'step %.4f ms' % d['ms_per_step'], 'prod %.0f' % d['production_mode']['value'],
      ' '.join('%s %.4f' % (n.replace('_kernel', ''), v['ms']) for n, v in k.items()), 'serial %.4f' % (r.get('serial_step_ms') or 0))"
  done
done
cp /tmp/libofdg_current.so $LIB
