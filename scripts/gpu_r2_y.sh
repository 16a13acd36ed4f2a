# round 2, call y: one ncu --set full capture of the bench's hot kernels on the final tree -> traffic.json (+ summaries)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"dcn_tc7_fwd|warp_fwd_staged" -c 20 -o gpurun_out/prof_final -f $CMD > gpurun_out/ncu_final.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_final.log | cut -c1-200
python scripts/traffic_from_report.py gpurun_out/prof_final.ncu-rep gpurun_out/traffic.json "$1" | cut -c1-900
python scripts/ncu_summary.py gpurun_out/prof_final.ncu-rep > gpurun_out/ncu_final_summary.txt 2>&1; grep -c "==" gpurun_out/ncu_final_summary.txt
