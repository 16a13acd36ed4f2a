mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['lsu_bound'], d['e2e'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 exit $?"; tail -3 gpurun_out/bench_cfg5.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg5.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
for m in bf16_tc fp32; do
timeout 600 python bench.py --workload cfg3 --math $m --steps 20 --warmup 5 > gpurun_out/bench_cfg3_${m}_n1.json 2> gpurun_out/bench_cfg3_${m}_n1.err; echo "cfg3 $m exit $?"; tail -3 gpurun_out/bench_cfg3_${m}_n1.err; cat gpurun_out/bench_cfg3_${m}_n1.json | cut -c1-600
done
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
