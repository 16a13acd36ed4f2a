# round 2, call p: TMA-staged warp -- parity (all warp tests), A/B against the L1 kernel, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" -k "warp or hot_path or smoke" > gpurun_out/pytest_warp.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_warp.log
timeout 300 python scripts/warp_ab.py staged_v1 > gpurun_out/warp_ab_staged.jsonl 2> gpurun_out/warp_ab_staged.err; echo "ab exit $?"; cat gpurun_out/warp_ab_staged.jsonl; tail -3 gpurun_out/warp_ab_staged.err
