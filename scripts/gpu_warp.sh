mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "warp or hot_path or smoke" > gpurun_out/t_warp.log 2>&1; echo "warp exit $?" >> gpurun_out/t_warp.log; tail -5 gpurun_out/t_warp.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; tail -2 gpurun_out/bench_w.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_w.json').read().strip().splitlines()[-1]); w=d['roofline_warp']; print('warp bf16', w['ms_per_launch'], w['frac'], 'f32', w['f32']['frac'], 'blend', d.get('roofline_blend')); print(d['ms_per_step'], d['roofline']['ms_per_launch'])"
