# round 2, call s: full GPU suite on the staged-warp tree, N=1 bench line, CPU full-frame validation of the stripe extrapolation
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); w=d['roofline_warp']; print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], 'warp', w['ms_per_launch'], w['frac'], w['f32'], w['in_step'], w.get('model_like_flow'), 'e2e', d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 900 python scripts/cpu_full_frame.py > gpurun_out/cpu_full_frame.json 2> gpurun_out/cpu_full_frame.err; echo "cpu exit $?"; cat gpurun_out/cpu_full_frame.json
