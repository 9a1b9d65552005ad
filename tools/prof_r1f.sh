python tools/prof_step.py > gpurun_out/plain_r1f.log 2>&1 &&
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:bn_bwd_reduce_k|bn_bwd_apply_k' -s 6 -c 2 -o gpurun_out/prof_bnbwd_pool_r1f -f python tools/prof_step.py > gpurun_out/ncu_r1f.log 2>&1
ls -la gpurun_out
