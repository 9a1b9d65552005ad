set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k "head or SpectralUNET or spectral" > gpurun_out/pytest_r2w.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2w.log
tail -n 3 gpurun_out/pytest_r2w.log
timeout 200 python bench.py --model SpectralUNET --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bd_spec_r2w.json > gpurun_out/bench_spec_r2w.json 2> gpurun_out/bench_spec_r2w.err
tail -n 3 gpurun_out/bench_spec_r2w.err
