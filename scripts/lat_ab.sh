for v in 4 1 2 8; do
  SVAE_LAT_BLOCKS_PER_SM=$v SVAE_TRACE=1 SVAE_MULTI=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-generation 2> /tmp/t.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('per_sm=$v ms/step %.3f'%d['ms_per_step'], [ (k['kernel'], round(k['ms_per_step'],3)) for k in d['kernels'] if k['kernel']=='skinny_fc'])"
  grep "TRACE skinny" /tmp/t.err | awk '{k=$4" "$5" "$7; n[k]++; s[k]+=substr($11,4)} END {for (k in n) printf "   %s n=%d avg=%.1f us\n", k, n[k], 1000*s[k]/n[k]}' | sort
done
