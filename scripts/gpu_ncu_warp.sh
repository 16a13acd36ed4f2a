mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"warp_fwd_fast" -s 6 -c 2 -o gpurun_out/prof_warp -f $CMD > gpurun_out/ncu_warp.log 2>&1
echo "ncu exit $?"
