# Round-2 evidence run (one B200): ncu launch list of one graph-replayed step with per-launch DRAM traffic, and
# `ncu --set full` captures of the kernels named in VERDICT r1 (weight gradient, batch-norm kernels, heads, conv).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation"
$CMD > gpurun_out/plain_r2.log 2>&1 || { tail -20 gpurun_out/plain_r2.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -s 2400 -c 1167 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_list_r2.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tc2_wgrad_kernel -s 170 -c 10 -o gpurun_out/prof_r2_wgrad $CMD > gpurun_out/ncu_full_wgrad.log 2>&1
echo "wgrad rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'bn_bwd_reduce_v8|bn_bwd_apply_v8|bn_act_fwd_v8' -s 450 -c 12 -o gpurun_out/prof_r2_bn $CMD > gpurun_out/ncu_full_bn.log 2>&1
echo "bn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'heads_|lat_' -s 140 -c 10 -o gpurun_out/prof_r2_heads $CMD > gpurun_out/ncu_full_heads.log 2>&1
echo "heads rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tc2_conv_kernel -s 360 -c 12 -o gpurun_out/prof_r2_conv $CMD > gpurun_out/ncu_full_conv.log 2>&1
echo "conv rc=$?"
ls -la gpurun_out/*.ncu-rep
