mkdir -p gpurun_out
python scripts/warp_one.py model_like staged > gpurun_out/warp_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"warp_fwd_staged" -s 4 -c 1 -o gpurun_out/prof_warp_staged -f python scripts/warp_one.py model_like staged > gpurun_out/ncu_warp_staged.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_warp_staged.log
ncu -i gpurun_out/prof_warp_staged.ncu-rep --page details > gpurun_out/ncu_warp_staged_details.txt 2>&1
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|Issue Slots Busy|No Eligible|Stall|L1/TEX Hit|L2 Hit|DRAM Throughput|Memory Throughput|Executed Ipc|Dynamic Shared|Block Limit" gpurun_out/ncu_warp_staged_details.txt | head -40
