mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "fp16_offsets or smoke" 2>&1 | tail -3
for w in cfg4 cfg4_iid cfg1 cfg5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit $?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); w=d['roofline_warp']; print('$w', d['value'], d['ms_per_step'], 'dcn', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'warp', w['ms_per_launch'], w['frac'], w['f32']['frac'], 'e2e', d['e2e']['value'])"
done
