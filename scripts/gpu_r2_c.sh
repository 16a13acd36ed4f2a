# round 2, call C: v7 geometry prefetch A/B (role counters + step time), then one ncu --set full capture of v7
mkdir -p gpurun_out
for v in default nopf; do
  if [ $v = default ]; then unset VFI_B200_LIB; else export VFI_B200_LIB=$PWD/video-frame-interpolation_b200/variants/libvfi_$v.so; fi
  echo "== $v"
  timeout 300 python scripts/dcn_debug7.py 2>&1 | tail -8
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
done
unset VFI_B200_LIB
timeout 300 python scripts/dcn_ab.py > gpurun_out/dcn_ab2.log 2>&1; echo "dcn_ab exit $?"; grep -A3 mismatches gpurun_out/dcn_ab2.log | head
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dcn_tc7_fwd" -s 6 -c 1 -o gpurun_out/prof_v7 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
