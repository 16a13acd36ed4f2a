mkdir -p gpurun_out
timeout 100 python scripts/wg_debug.py 2 256 256 2>&1 | head -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "weight_grad_tensor_core or umma or tensor_core or fused or hot_path or variants" > gpurun_out/t_wg.log 2>&1; echo "wg exit $?" >> gpurun_out/t_wg.log; tail -3 gpurun_out/t_wg.log | cut -c1-220
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['ms_per_launch'])"
for m in bf16_tc; do timeout 120 python scripts/train_step_bench.py --math $m > gpurun_out/train_cfg3_$m.json 2> gpurun_out/train_cfg3_$m.err; echo "$m exit $?"; tail -1 gpurun_out/train_cfg3_$m.json | cut -c150-420; tail -2 gpurun_out/train_cfg3_$m.err; done
