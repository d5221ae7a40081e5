for b in 16 32 64 128; do python bench.py --no-cpu --no-layer --no-other-configs --e2e-steps 3 --steps 400 --batch $b --no-attribution 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $b', 'samples/s %.0f'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'prod %.0f'%d['production_mode']['value'])"; done
