mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -k "fused or hot_path or variants or tensor_core_path" 2>&1 | tail -4
for l in nchw channels_last nchw channels_last; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --conv27-layout $l 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
done
