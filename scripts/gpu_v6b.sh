mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "umma or tensor_core or fused or hot_path or variants" > gpurun_out/t_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/t_tc.log; tail -5 gpurun_out/t_tc.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
tail -3 gpurun_out/bench_v6.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_v6.json').read().strip().splitlines()[-1]); print('v6', d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['clocks'])"
timeout 300 python scripts/dcn_debug.py 2>&1 | grep -v Warning | grep "producers\|mma \|epilogue\|geometry\|tiles per"
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
