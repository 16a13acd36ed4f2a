mkdir -p gpurun_out
for v in default p2; do
  if [ $v = default ]; then unset VFI_B200_LIB; export V7_PARTS=3; else export VFI_B200_LIB=$PWD/video-frame-interpolation_b200/variants/libvfi_$v.so; export V7_PARTS=2; fi
  echo "== $v"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
  VFI_DCN_EXPERIMENT=0 timeout 300 python scripts/dcn_debug7.py 2>&1 | tail -8
done
unset VFI_B200_LIB; export V7_PARTS=3
VFI_DCN_EXPERIMENT=13 timeout 300 python scripts/dcn_debug7.py 2>&1 | tail -8
timeout 300 python scripts/dcn_ab.py > gpurun_out/dcn_ab7.log 2>&1; echo "dcn_ab exit $?"; grep -A3 mismatches gpurun_out/dcn_ab7.log | head
