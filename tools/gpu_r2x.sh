set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "conv3x3" > gpurun_out/pytest_r2x.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2x.log
tail -n 3 gpurun_out/pytest_r2x.log
HPRI_HALO_A_STAGES=3 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k "conv3x3 or train_step_parity_vs_oracle" > gpurun_out/pytest_r2x3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2x3.log
tail -n 3 gpurun_out/pytest_r2x3.log
python tools/ab_step.py 6 10 default,halo_a_stages3 > gpurun_out/ab_r2x.jsonl 2> gpurun_out/ab_r2x.err
cat gpurun_out/ab_r2x.jsonl; tail -n 3 gpurun_out/ab_r2x.err
