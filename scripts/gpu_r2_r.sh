# round 2, call r: launch list of the cfg3 bf16 training step (per-kernel time shares), full GPU suite on the staged-warp tree
mkdir -p gpurun_out
CMD="python bench.py --workload cfg3 --math bf16_tc --steps 2 --warmup 3"
$CMD > gpurun_out/cfg3_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/ncu_train.log 2>&1
echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/train_launches.csv > gpurun_out/train_launches.txt 2>&1; head -40 gpurun_out/train_launches.txt
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
