mkdir -p gpurun_out
# tensor-core plumbing first, each under its own timeout so a protocol bug cannot hang the box
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "umma_selftest" > gpurun_out/t_umma.log 2>&1; echo "umma exit $?" >> gpurun_out/t_umma.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "tensor_core" > gpurun_out/t_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/t_tc.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_umma_selftest_matches_matmul -k "not tensor_core" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -4 gpurun_out/t_umma.log; tail -4 gpurun_out/t_tc.log; tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench.err; head -c 2500 gpurun_out/bench.json
