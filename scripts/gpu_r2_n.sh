mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" -k "column_gradient or data_grads or config3 or training or smoke or fusion_block" > gpurun_out/pytest_bwd.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_bwd.log
timeout 600 python bench.py --workload cfg3 --math bf16_tc --steps 20 --warmup 5 > gpurun_out/bench_cfg3_gcol.json 2> gpurun_out/bench_cfg3_gcol.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_gcol.json').read().strip().splitlines()[-1]); print('cfg3', d['ms_per_step'], d['value'])"
timeout 300 python scripts/bwd_breakdown.py > gpurun_out/bwd_breakdown.json 2>&1; tail -5 gpurun_out/bwd_breakdown.json
