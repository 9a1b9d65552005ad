# round-1 (h) evidence set (CTA-pair conv kernels, weight gradients on the side stream): one CubeNET-64 training step (tools/prof_step.py, batch 2, 238x608x968), profiled range only.
# Every ncu command runs after the same command line exited 0 without ncu (first line).
set -x
python tools/prof_step.py > gpurun_out/plain_r1h.log 2>&1 &&
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_r1h.csv python tools/prof_step.py > gpurun_out/ncu_r1h_a.log 2>&1
# source-level captures: halo<64> first_conv + inc2, halo<128> down1.c1 + down1.c2
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 4 -o gpurun_out/prof_halo_r1h -f python tools/prof_step.py > gpurun_out/ncu_r1h_b.log 2>&1
# wgrad N=64 (up4.c2, up4.c1) and the last launch of the step (first_conv wgrad) are igemm_kernel<64,4,1>
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wgrad3x3_halo|igemm_kernel.*256.*2.*1>' -c 6 -o gpurun_out/prof_wgrad64_r1h -f python tools/prof_step.py > gpurun_out/ncu_r1h_c.log 2>&1
# memory-bound kernels: ingest, first BN apply / backward launches
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:ingest|bn_bwd|bn_relu_apply|head_fwd|bce' -c 12 -o gpurun_out/prof_elem_r1h -f python tools/prof_step.py > gpurun_out/ncu_r1h_d.log 2>&1
ls -la gpurun_out
