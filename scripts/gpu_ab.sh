mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "tensor_core or fused or hot_path" > gpurun_out/t_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/t_tc.log; tail -3 gpurun_out/t_tc.log
for k in v4 v5; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --dcn-kernel $k > gpurun_out/bench_$k.json 2> gpurun_out/bench_$k.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_$k.json').read().strip().splitlines()[-1]); print('$k', d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
