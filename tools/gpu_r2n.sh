set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "pack or table or convT" > gpurun_out/pytest_r2n.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2n.log
tail -n 4 gpurun_out/pytest_r2n.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_r2n.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2n.log; tail -n 3 gpurun_out/smoke_r2n.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --breakdown gpurun_out/bd_r2n.json > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err
tail -n 3 gpurun_out/bench_r2n.err
