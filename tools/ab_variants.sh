#!/bin/bash
# A/B of prebuilt library variants on one GPU box: tools/ab_variants.sh variants/libofdg_a.so variants/libofdg_b.so ...
# (each variant is copied over csrc/libofdg.so, bench.py runs without the CPU leg; one summary line per variant)
LIB=optical-flow-2d-data-generation_b200/csrc/libofdg.so
cp $LIB /tmp/libofdg_current.so
for v in "$@" /tmp/libofdg_current.so; do
  cp "$v" $LIB
  for rep in 1 2; do
    python bench.py --no-cpu --e2e-steps 3 --steps 600 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('$v', 'samples/s %.0f' % d['value'], 'prod %.0f' % d['production_mode']['value'], 'prep %.4f raster %.4f shade %.4f frac %.3f' % (r['bg_prep_ms'], r['raster_ms'] or 0, r['kernel_ms'], r['frac']))"
  done
done
cp /tmp/libofdg_current.so $LIB
