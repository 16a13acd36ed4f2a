# round 2, call E: straight-line geometry; diagnostics (VFI_DCN_EXPERIMENT) that remove one role's work at a time
mkdir -p gpurun_out
for v in default nopf; do
  if [ $v = default ]; then unset VFI_B200_LIB; else export VFI_B200_LIB=$PWD/video-frame-interpolation_b200/variants/libvfi_$v.so; fi
  echo "== $v"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
done
unset VFI_B200_LIB
for e in 0 1 4 8 13 2 15; do
  echo "== experiment $e"
  VFI_DCN_EXPERIMENT=$e timeout 300 python scripts/dcn_debug7.py 2>&1 | tail -8
done
timeout 300 python scripts/dcn_ab.py > gpurun_out/dcn_ab4.log 2>&1; echo "dcn_ab exit $?"; grep -A3 mismatches gpurun_out/dcn_ab4.log | head
