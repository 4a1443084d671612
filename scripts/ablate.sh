# timing-only ablations of the train step (SVAE_ABLATE makes results wrong on purpose): how long is the chain alone?
for a in 0 1 2 3; do
  SVAE_ABLATE=$a python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-generation 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ablate=$a ms_per_step=%.3f'%d['ms_per_step'])"
done
