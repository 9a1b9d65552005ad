set -x
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -x -p no:cacheprovider -k "wgrad or convT or linear or train_step_parity_vs_oracle or SpectralUNET" > gpurun_out/pytest_r2y.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2y.log
tail -n 6 gpurun_out/pytest_r2y.log
timeout 200 python bench.py --model SpectralUNET --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bd_spec_r2y.json > gpurun_out/bench_spec_r2y.json 2> gpurun_out/bench_spec_r2y.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --breakdown gpurun_out/bd_r2y.json > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err
tail -n 3 gpurun_out/bench_spec_r2y.err gpurun_out/bench_r2y.err
