mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "test_dcn_fused_split_input_and_conv27 or test_warp_into_tail or test_dcn_tensor_core_path_fp16" > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/memcheck.log; grep "ERROR SUMMARY\|Invalid\|passed\|failed" gpurun_out/memcheck.log | head -10
