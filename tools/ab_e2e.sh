#!/bin/bash
# A/B of prebuilt library variants on the host-blob (e2e) leg: tools/ab_e2e.sh variants/a.so ... ("-" = the library in place)
LIB=optical-flow-2d-data-generation_b200/csrc/libofdg.so
cp $LIB /tmp/libofdg_current.so
for v in "$@"; do
  if [ "$v" = "-" ]; then cp /tmp/libofdg_current.so $LIB; else cp "$v" $LIB; fi
  python bench.py --no-cpu --no-layer --no-other-configs --no-attribution --e2e-steps 40 --steps 100 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']
print('[$v]', 'e2e %.0f samples/s' % e['value'], 'dram %.0f GB/s implied' % e['host_dram_gbs_implied'], 'device %.0f' % d['value'])"
done
cp /tmp/libofdg_current.so $LIB
