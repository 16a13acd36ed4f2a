mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dcn_tc" -s 6 -c 1 -o gpurun_out/prof_tc -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/plain.log | cut -c1-600
