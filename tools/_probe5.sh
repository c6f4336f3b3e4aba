mkdir -p gpurun_out/p5
./tools/micro/ffma_bank_bench > gpurun_out/p5/ffma_bank.log 2>&1
ncu --metrics smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p5/ffma_bank_ncu.csv ./tools/micro/ffma_bank_bench > /dev/null 2>&1
cat gpurun_out/p5/ffma_bank.log
python -m pytest tests -x -q -m gpu > gpurun_out/p5/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p5/pytest_gpu.log
