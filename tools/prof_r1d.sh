# round-1 (d): source-level captures of the halo conv kernel (inc2 64->64 and down1.c2 128->128 forward)
set -x
python tools/prof_step.py > gpurun_out/plain_r1d.log 2>&1 &&
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:halo -s 1 -c 1 -o gpurun_out/prof_halo64_inc2_r1d -f python tools/prof_step.py > gpurun_out/ncu_r1d_a.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:halo -s 3 -c 1 -o gpurun_out/prof_halo128_down1c2_r1d -f python tools/prof_step.py > gpurun_out/ncu_r1d_b.log 2>&1
ls -la gpurun_out
