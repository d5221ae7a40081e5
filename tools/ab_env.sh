#!/bin/bash
# A/B of run-time options of one library on one GPU box: tools/ab_env.sh "VAR=1 OTHER=2" "VAR=0" ...
# (bench.py without the CPU leg under each environment; one summary line per setting; "-" = defaults)
for v in "$@"; do
  e="$v"; [ "$v" = "-" ] && e=""
  env $e python bench.py --no-cpu --e2e-steps 3 --steps 600 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
k = r.get('kernels') or {}
print('[$v]', 'samples/s %.0f' % d['value'], 'prod %.0f' % d['production_mode']['value'], ' '.join('%s %.4f' % (n.replace('_kernel', ''), x['ms']) for n, x in k.items()), 'frac %.3f' % r['frac'])"
done
