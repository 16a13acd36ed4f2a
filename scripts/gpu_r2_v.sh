mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" -k "fused_training or fused_blocks or cuda_graph" > gpurun_out/pytest_fused.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_fused.log | cut -c1-300
for g in "--fused-blocks" "--fused-blocks --cuda-graph"; do
t=$(echo $g | tr -d ' -')
timeout 90 python bench.py --workload cfg3 --math bf16_tc --steps 20 --warmup 5 $g > gpurun_out/bench_cfg3_$t.json 2> gpurun_out/bench_cfg3_$t.err; echo "cfg3 '$g' exit $?"; tail -2 gpurun_out/bench_cfg3_$t.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_$t.json').read().strip().splitlines()[-1]); print('cfg3 $g', d['ms_per_step'], d['value'], d['gpu_launches'], d['gradients_finite'])"
done
