for v in "SVAE_ABLATE=0" "SVAE_ABLATE=1" "SVAE_ABLATE=2" "SVAE_ABLATE=3" "SVAE_ABLATE=3 SVAE_FORK_MASK=0"; do
  echo "== $v"; env $v FLOOR_BS=2 timeout 200 python scripts/latency_floor.py 2>&1 | tail -1
done
