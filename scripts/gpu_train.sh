mkdir -p gpurun_out
for m in bf16_tc; do timeout 120 python scripts/train_step_bench.py --math $m > gpurun_out/train_cfg3_$m.json 2> gpurun_out/train_cfg3_$m.err; echo "$m exit $?"; tail -1 gpurun_out/train_cfg3_$m.json | cut -c150-400; tail -2 gpurun_out/train_cfg3_$m.err; done
CMD="timeout 120 python scripts/train_step_bench.py --math bf16_tc --steps 2 --warmup 1"
$CMD > gpurun_out/train_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches_tc.csv $CMD > gpurun_out/train_ncu.log 2>&1
echo "list exit $?"
