mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python scripts/model_bench.py --batch 1 4 --out gpurun_out/model_bench.jsonl > gpurun_out/model_bench.log 2>&1; echo "model_bench exit $?"; tail -20 gpurun_out/model_bench.log | cut -c1-400
