# round-1 (c) captures: one CubeNET-64 training step bracketed by cudaProfilerStart/Stop (tools/prof_step.py)
set -x
python tools/prof_step.py > gpurun_out/plain_r1c.log 2>&1 &&
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/prof_step.py > gpurun_out/ncu_r1c_a.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:halo -s 1 -c 1 -o gpurun_out/prof_halo_inc2_r1c -f python tools/prof_step.py > gpurun_out/ncu_r1c_c.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:igemm_kernel.*64.*4.*1>' -s 1 -c 1 -o gpurun_out/prof_wgrad_up4c1_r1c -f python tools/prof_step.py > gpurun_out/ncu_r1c_d.log 2>&1
ls -la gpurun_out
tail -2 gpurun_out/ncu_r1c_*.log
