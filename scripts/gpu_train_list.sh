mkdir -p gpurun_out
CMD="python scripts/train_step_bench.py --steps 2 --warmup 1"
$CMD > gpurun_out/train_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1
echo "list exit $?"
