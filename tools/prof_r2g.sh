# round-2 (g) evidence set.  Every ncu command follows a plain run of the same command line that exited 0.
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
# 1. launch list of the bench command itself
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_bench_r2g.log 2>&1 &&
timeout 400 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_r2g.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r2g_1.log 2>&1
# 2. one CubeNET training step: every launch with time / DRAM bytes / tensor-pipe share (profiled range only)
python tools/prof_step.py > gpurun_out/plain_step_r2g.log 2>&1 &&
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_step_r2g.csv python tools/prof_step.py > gpurun_out/ncu_r2g_2.log 2>&1
# 3. --set full: the memory-bound family of that step (ingest, BN apply / backward, head, loss, pack / unpack, sums)
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:ingest|bn_|head_fwd|bce_k|pack_conv|colsum|sum_f32' -c 60 -o gpurun_out/prof_elem_r2g -f python tools/prof_step.py > gpurun_out/ncu_r2g_3.log 2>&1
# 4. --set full: tensor kernels -- first six halo convs, first six weight gradients
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:conv3x3_halo' -c 6 -o gpurun_out/prof_halo_r2g -f python tools/prof_step.py > gpurun_out/ncu_r2g_4.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wgrad3x3_halo|igemm_kernel<.*1>' -c 6 -o gpurun_out/prof_wgrad_r2g -f python tools/prof_step.py > gpurun_out/ncu_r2g_5.log 2>&1
# 5. SpectralUNET step (one 608 x 700 image): the GEMM chain and its strided BatchNorm kernels
python tools/prof_step.py SpectralUNET > gpurun_out/plain_spectral_r2g.log 2>&1 &&
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:igemm_kernel|bn_|head_fwd' -c 14 -o gpurun_out/prof_spectral_r2g -f python tools/prof_step.py SpectralUNET > gpurun_out/ncu_r2g_6.log 2>&1
# 6. kernels outside the training step
python tools/prof_step.py misc > gpurun_out/plain_misc_r2g.log 2>&1 &&
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:adam_k|pr_hist_k|upsample2|mul16' -c 8 -o gpurun_out/prof_misc_r2g -f python tools/prof_step.py misc > gpurun_out/ncu_r2g_7.log 2>&1
ls -la gpurun_out | tail -20
tail -2 gpurun_out/ncu_r2g_*.log
