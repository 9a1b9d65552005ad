# round-2 (c): re-run of the touched tests, then the bench with and without the next-batch ingest prefetch
set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_trainer_gpu.py -m gpu -q -x -p no:cacheprovider -k "bn_ or bce or prefetch or param_grads or batch_tables or adam or generic_criterion or roundtrip or head" > gpurun_out/pytest_r2c.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2c.log
tail -5 gpurun_out/pytest_r2c.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --breakdown gpurun_out/bd_r2c.json > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
HPRI_INGEST_PREFETCH=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2c_nopf.json 2> gpurun_out/bench_r2c_nopf.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2c_pf2.json 2>> gpurun_out/bench_r2c.err
tail -3 gpurun_out/bench_r2c.err
