mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" -k "cuda_graph or fusion_block or records" > gpurun_out/pytest_graph.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_graph.log
for g in "" "--cuda-graph"; do
timeout 600 python bench.py --workload cfg3 --math bf16_tc --steps 20 --warmup 5 $g > gpurun_out/bench_cfg3_graph$g.json 2> gpurun_out/bench_cfg3_graph$g.err; echo "cfg3 '$g' exit $?"; tail -2 gpurun_out/bench_cfg3_graph$g.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_graph$g.json').read().strip().splitlines()[-1]); print('cfg3 $g', d['ms_per_step'], d['value'], d['gpu_launches'], d['config'].get('launch'))"
done
