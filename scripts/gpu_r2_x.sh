# round 2, final call: full GPU suite, smoke, short N=1 bench line (GPU budget nearly spent: keep it tight)
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -p no:cacheprovider -W "ignore::RuntimeWarning" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -8 gpurun_out/smoke.log | cut -c1-300
