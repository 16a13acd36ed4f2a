mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 exit $?" >> gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 exit $?" >> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "refn2 exit $?" >> gpurun_out/bench_ref_n2.err
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_n1.err gpurun_out/bench_n2.err gpurun_out/bench_ref.err gpurun_out/bench_ref_n2.err
python - <<'PY'
import json
for f in ('bench_n1','bench_n2','bench_ref','bench_ref_n2'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), 'roof', (d.get('roofline') or {}).get('frac'), (d.get('roofline_warp') or {}).get('frac'), ((d.get('roofline_warp') or {}).get('f32') or {}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
