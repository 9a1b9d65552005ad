set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k "bn_ or head or train_step_parity or odd_sizes" > gpurun_out/pytest_r2q.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2q.log
tail -n 3 gpurun_out/pytest_r2q.log
python tools/ab_step.py 6 10 default,ascending_elementwise > gpurun_out/ab_r2q.jsonl 2> gpurun_out/ab_r2q.err
cat gpurun_out/ab_r2q.jsonl; tail -n 3 gpurun_out/ab_r2q.err
