# round 2, call B: first run of the v7 DCN forward -- bit-exactness against v6, timing, role counters, parity tests
mkdir -p gpurun_out
timeout 900 python scripts/dcn_ab.py --time --out gpurun_out/dcn_ab.json > gpurun_out/dcn_ab.log 2>&1; echo "dcn_ab exit $?"
tail -40 gpurun_out/dcn_ab.log
timeout 300 python scripts/dcn_debug7.py > gpurun_out/dcn_debug7.txt 2>&1; echo "dcn_debug7 exit $?"; cat gpurun_out/dcn_debug7.txt | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "dcn or hot_path or smoke or umma" > gpurun_out/pytest_dcn.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_dcn.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_v7.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac_of_burst_peak'])"
