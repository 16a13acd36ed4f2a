# A/B of kernel-tuning variants on one box: DCN ms per layer, same inputs
run() { VFI_B200_LIB="$2" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['ms_per_step'], d['roofline']['ms_per_launch'])"; }
run base ""
for v in $(ls video-frame-interpolation_b200/variants/*.so); do run $(basename $v .so) $PWD/$v; done
run base ""
