mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "weight_grad_tensor_core" > gpurun_out/t_wg.log 2>&1; echo "wg exit $?" >> gpurun_out/t_wg.log; tail -30 gpurun_out/t_wg.log | cut -c1-220
