set -x
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2z.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2z.log; tail -n 2 gpurun_out/smoke_r2z.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_r2z.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2z.log
tail -n 3 gpurun_out/pytest_r2z.log
