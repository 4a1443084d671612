for v in "X=1" "SVAE_SKIP_F32=0" "SVAE_PDL=0" "SVAE_CONV_2CTA=0"; do echo "== $v"; env $v timeout 120 python scripts/smoke_diag.py 2>&1 | tail -8; done
