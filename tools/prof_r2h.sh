# round-2 (h) evidence set.  Every ncu command follows a plain run of the same command line that exited 0.  Reports are
# flattened to CSV on the box (tools/ncu_summary.py) and the large .ncu-rep files removed: gpurun_out/ is capped at 64 MiB.
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
S=gpurun_out
# 1. launch list of the bench command itself (time, DRAM bytes, tensor-pipe share of every launch)
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $S/plain_bench_r2h.log 2>&1 &&
timeout 400 ncu --metrics $M --clock-control none -c 600 --csv --log-file $S/launches_bench_r2h.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $S/ncu_r2h_1.log 2>&1
# 2. --set full: the memory-bound family of one CubeNET training step
python tools/prof_step.py > $S/plain_step_r2h.log 2>&1 &&
timeout 400 ncu --profile-from-start off --set full --clock-control none -k 'regex:ingest|bn_|head_fwd|bce_k|pack_conv|colsum|sum_f32' -c 44 -o $S/prof_elem_r2h -f python tools/prof_step.py > $S/ncu_r2h_2.log 2>&1
python tools/ncu_summary.py $S/ncu_full_r2h_elementwise.csv $S/prof_elem_r2h.ncu-rep && rm -f $S/prof_elem_r2h.ncu-rep
# 3. --set full with source: tensor kernels -- first four halo convs, first four weight gradients
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:conv3x3_halo' -c 4 -o $S/prof_halo_r2h -f python tools/prof_step.py > $S/ncu_r2h_3.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wgrad3x3_halo|igemm_kernel<.*1>' -c 4 -o $S/prof_wgrad_r2h -f python tools/prof_step.py > $S/ncu_r2h_4.log 2>&1
python tools/ncu_summary.py $S/ncu_full_r2h_tensor.csv $S/prof_halo_r2h.ncu-rep $S/prof_wgrad_r2h.ncu-rep
# 4. SpectralUNET step (one 608 x 700 image): GEMM chain forward (tail K=240, down1 K=1664), backward (wgrad / dgrad), BatchNorm
python tools/prof_step.py SpectralUNET > $S/plain_spectral_r2h.log 2>&1 &&
timeout 300 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k 'regex:igemm_kernel' -c 2 -o $S/prof_spec_fwd_r2h -f python tools/prof_step.py SpectralUNET > $S/ncu_r2h_5.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k 'regex:igemm_kernel' -s 9 -c 4 -o $S/prof_spec_bwd_r2h -f python tools/prof_step.py SpectralUNET > $S/ncu_r2h_6.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:bn_|head_fwd' -c 5 -o $S/prof_spec_bn_r2h -f python tools/prof_step.py SpectralUNET > $S/ncu_r2h_7.log 2>&1
python tools/ncu_summary.py $S/ncu_full_r2h_spectral.csv $S/prof_spec_fwd_r2h.ncu-rep $S/prof_spec_bwd_r2h.ncu-rep $S/prof_spec_bn_r2h.ncu-rep && rm -f $S/prof_spec_*_r2h.ncu-rep
# 5. kernels outside the training step
python tools/prof_step.py misc > $S/plain_misc_r2h.log 2>&1 &&
timeout 200 ncu --profile-from-start off --set full --clock-control none -k 'regex:adam_k|pr_hist_k|upsample2|mul16' -c 8 -o $S/prof_misc_r2h -f python tools/prof_step.py misc > $S/ncu_r2h_8.log 2>&1
python tools/ncu_summary.py $S/ncu_full_r2h_misc.csv $S/prof_misc_r2h.ncu-rep && rm -f $S/prof_misc_r2h.ncu-rep
du -sh $S; ls -la $S
for f in $S/ncu_r2h_*.log; do echo == $f; tail -n 2 $f; done
