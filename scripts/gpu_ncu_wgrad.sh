mkdir -p gpurun_out
CMD="timeout 150 python scripts/bwd_breakdown.py"
$CMD > gpurun_out/plain_bwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"wgrad" -c 2 -o gpurun_out/prof_wgrad -f $CMD > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu exit $?"; tail -1 gpurun_out/plain_bwd.log | cut -c1-600
