# launch list + full capture of the two hot kernels for the bench command (run under gpurun)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"dcn_tc_fwd|warp_fwd" -s 8 -c 6 -o gpurun_out/prof_full -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"; tail -2 gpurun_out/plain.log | cut -c1-400
