# round 2, call D: warp-role layout A/B (helpers first / last) with batched tail loads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw,temperature.gpu --format=csv
for v in default hl; do
  if [ $v = default ]; then unset VFI_B200_LIB; export V7_LAYOUT=1; else export VFI_B200_LIB=$PWD/video-frame-interpolation_b200/variants/libvfi_$v.so; export V7_LAYOUT=0; fi
  echo "== $v"
  timeout 300 python scripts/dcn_debug7.py 2>&1 | tail -8
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['clocks'])"
done
unset VFI_B200_LIB
VFI_DCN_KERNEL=v6 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench v6', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['clocks'])"
timeout 300 python scripts/dcn_ab.py > gpurun_out/dcn_ab3.log 2>&1; echo "dcn_ab exit $?"; grep -A3 mismatches gpurun_out/dcn_ab3.log | head
