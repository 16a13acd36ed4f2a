mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dcn_tc7_fwd" -s 6 -c 3 -o gpurun_out/prof_v7b -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
VFI_DCN_DEBUG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench with DBG kernel', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
