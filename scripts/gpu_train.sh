mkdir -p gpurun_out
timeout 600 python scripts/train_step_bench.py > gpurun_out/train_cfg3.json 2> gpurun_out/train_cfg3.err; echo "ours exit $?"; tail -1 gpurun_out/train_cfg3.json; tail -3 gpurun_out/train_cfg3.err
timeout 600 python scripts/train_step_bench.py --stock > gpurun_out/train_cfg3_stock.json 2> gpurun_out/train_cfg3_stock.err; echo "stock exit $?"; tail -1 gpurun_out/train_cfg3_stock.json; tail -3 gpurun_out/train_cfg3_stock.err
