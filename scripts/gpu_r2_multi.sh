# usage: gpu_r2_multi.sh N   -- multi-GPU evidence on one box: 2-GPU parity test, cfg3 (training, all-reduce), cfg2, cfg5, PCIe probe
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ $N = 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -W "ignore::RuntimeWarning" > gpurun_out/pytest_gpu_multi_n2.log 2>&1; echo "pytest multi exit $?" >> gpurun_out/pytest_gpu_multi_n2.log; tail -4 gpurun_out/pytest_gpu_multi_n2.log
fi
timeout 600 $TR --master-port 29511 bench.py --gpus $N --workload cfg3 --math bf16_tc --steps 20 --warmup 5 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err; echo "cfg3 overlap exit $?"; tail -1 gpurun_out/bench_cfg3_n$N.json | cut -c1-160; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_n$N.json').read().strip().splitlines()[-1]); print('cfg3 overlap', d['ms_per_step'], d['value'], d['allreduce_exposed_us'])"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --workload cfg3 --math bf16_tc --steps 20 --warmup 5 --no-overlap > gpurun_out/bench_cfg3_noov_n$N.json 2> gpurun_out/bench_cfg3_noov_n$N.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_noov_n$N.json').read().strip().splitlines()[-1]); print('cfg3 no overlap', d['ms_per_step'], d['value'], d['allreduce_exposed_us'])"
timeout 600 $TR --master-port 29516 bench.py --gpus $N --workload cfg3 --math bf16_tc --steps 20 --warmup 5 --cuda-graph > gpurun_out/bench_cfg3_graph_n$N.json 2> gpurun_out/bench_cfg3_graph_n$N.err; echo "cfg3 graph exit $?"; tail -2 gpurun_out/bench_cfg3_graph_n$N.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_graph_n$N.json').read().strip().splitlines()[-1]); print('cfg3 cuda graph', d['ms_per_step'], d['value'])"
[ -n "$ONLY_CFG3" ] && exit 0
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "cfg2 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_gbs_this_rank'], d['e2e']['cpu_affinity'])"
timeout 900 $TR --master-port 29514 bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_n$N.json 2> gpurun_out/bench_cfg5_n$N.err; echo "cfg5 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg5_n$N.json').read().strip().splitlines()[-1]); print('cfg5', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 300 $TR --master-port 29515 scripts/pcie_probe.py > gpurun_out/pcie_probe_n$N.txt 2>&1; tail -4 gpurun_out/pcie_probe_n$N.txt
