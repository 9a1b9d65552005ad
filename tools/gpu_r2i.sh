set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k "head or bn_ or deterministic or linear or SpectralUNET or param_grads" > gpurun_out/pytest_r2i.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2i.log
tail -6 gpurun_out/pytest_r2i.log
HPRI_SPECTRAL_WGRAD_TILE=128 timeout 200 python bench.py --model SpectralUNET --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_spec_t128_r2i.json 2> gpurun_out/bench_spec_t128_r2i.err
HPRI_SPECTRAL_WGRAD_TILE=256 timeout 200 python bench.py --model SpectralUNET --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bd_spec_t256_r2i.json > gpurun_out/bench_spec_t256_r2i.json 2> gpurun_out/bench_spec_t256_r2i.err
timeout 400 python bench.py --breakdown gpurun_out/bd_r2i.json > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
for f in gpurun_out/*_r2i.err; do echo == $f; grep -v "^$" $f | tail -n 5; done
