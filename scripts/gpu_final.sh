# round-end evidence: tests, bench (both arms), launch list, one ncu --set full capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"dcn_tc6|warp_fwd_fast" -s 8 -c 4 -o gpurun_out/prof_full -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
