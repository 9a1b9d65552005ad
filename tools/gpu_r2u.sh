set -x
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k "bn_ or SpectralUNET or spectral or head" > gpurun_out/pytest_r2u.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2u.log
tail -n 4 gpurun_out/pytest_r2u.log
timeout 200 python bench.py --model SpectralUNET --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bd_spec_r2u.json > gpurun_out/bench_spec_r2u.json 2> gpurun_out/bench_spec_r2u.err
tail -n 3 gpurun_out/bench_spec_r2u.err
