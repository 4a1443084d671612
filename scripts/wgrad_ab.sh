# weight-gradient kernel per layer shape under different cuts (event-bracketed launches, SVAE_TRACE): default = cost model
mkdir -p gpurun_out
for v in "" "SVAE_WGRAD_MODEL=0" "$@"; do
  echo "=== [$v]"
  env $v SVAE_TRACE=1 python scripts/bench_ops.py 2>&1 | grep "TRACE wgrad" | awk '{k=$3" "$4" "$5" "$6" "$7" "$9" "$10; n[k]++; v[k]=$11} END {for (k in v) print k, v[k]}' | sort
done
