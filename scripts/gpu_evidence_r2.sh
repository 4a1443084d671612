# Round-2 evidence run on one B200.  Everything that comes back is text (the .ncu-rep files stay on the box: four of them
# exceed the 64 MiB that gpurun copies back): bench lines, the ncu launch list of one graph-replayed step with per-launch DRAM
# bytes, and the raw pages of `ncu --set full` captures of the kernels VERDICT r1 asked for.
mkdir -p gpurun_out
TAG=${1:-r2}
python bench.py > gpurun_out/${TAG}_bench_final.json 2> gpurun_out/${TAG}_bench_final.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/${TAG}_bench_final.json | head -14
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-generation"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -s 2550 -c 842 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "list rc=$?"
cap() {   # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/prof_$1 $CMD > gpurun_out/${TAG}_ncu_full_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_$1.csv 2>/dev/null
  ls -la /tmp/prof_$1.ncu-rep | awk '{print $5}'
}
cap wgrad 'tc2_wgrad_kernel' 60 10
cap bn 'bn_bwd_reduce_v8_kernel|bn_bwd_apply_v8_kernel|bn_act_fwd_v8_kernel' 300 12
cap conv 'tc2_conv_kernel' 150 12
cap multi 'multi_kernel' 60 14
cp /tmp/prof_wgrad.ncu-rep gpurun_out/${TAG}_prof_wgrad.ncu-rep 2>/dev/null
ls -la gpurun_out | head -40
