mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dcn_tc_fwd|warp_fwd" -s 4 -c 4 -o gpurun_out/prof_tc -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
