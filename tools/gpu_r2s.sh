set -x
timeout 500 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_dp_r2s.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_dp_r2s.log
tail -n 30 gpurun_out/pytest_dp_r2s.log
