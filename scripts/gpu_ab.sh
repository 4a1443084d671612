# A/B of environment switches on one B200: `scripts/gpu_ab.sh "A=1" "B=2 C=3" ...` runs the short bench once per variant
# (first the default build), prints ms/step, img/s and launches per step of each.
mkdir -p gpurun_out
i=0
for v in "" "$@"; do
  out=gpurun_out/ab_$i.json
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-generation > $out 2> gpurun_out/ab_$i.err || tail -c 600 gpurun_out/ab_$i.err
  python - "$out" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("variant [%s]: %.3f ms/step  %.0f img/s  e2e %.0f  launches %s" % (sys.argv[2], d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("gpu_launches")))
except Exception as e:
    print("variant [%s]: failed (%s)" % (sys.argv[2], e))
PY
  i=$((i+1))
done
