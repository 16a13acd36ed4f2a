# round 2, call A: state of the round-1 build on this round's box + the gather-only microbenchmark + raw ncu metrics
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 scripts/microbench/gather_floor > gpurun_out/gather_floor.jsonl 2> gpurun_out/gather_floor.err; echo "gather_floor exit $?"
timeout 300 python scripts/dcn_debug.py > gpurun_out/dcn_debug.txt 2>&1; echo "dcn_debug exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
for w in cfg4 cfg4_iid cfg5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit $?"
done
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"dcn_tc6_fwd" -s 6 -c 1 -o gpurun_out/prof_tc -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
cat gpurun_out/gather_floor.jsonl
cat gpurun_out/dcn_debug.txt
python -c "
import json
for w in ['n1','cfg4','cfg4_iid','cfg5']:
    try:
        d=json.loads(open(f'gpurun_out/bench_{w}.json').read().strip().splitlines()[-1]); print(w, d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline_warp']['frac'], (d.get('e2e') or {}).get('value'))
    except Exception as e: print(w, 'failed', e)
"
