#!/bin/bash
# Host-blob transport A/B at N ranks on one box: bash tools/exp_transport_n.sh N
# (bytes + host widening against plain float blobs; everything else of the bench switched off)
N=${1:-8}
for t in u8 f32 auto; do
  if [ $t = auto ]; then unset OFDG_TRANSPORT; else export OFDG_TRANSPORT=$t; fi
  python bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-layer --no-other-configs --no-attribution --no-gather --e2e-steps 40 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']
print('N=$N transport=$t', 'device %.0f samples/s' % d['value'], 'e2e %.0f samples/s' % e['value'], 'd2h %.0f MB/step' % (e['d2h_bytes_per_step'] / 1e6), 'host dram implied %.0f GB/s' % e['host_dram_gbs_implied'], 'prod %.0f' % d['production_mode']['value'])"
done
