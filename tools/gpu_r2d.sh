set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k "bn_contiguous or prefetch" > gpurun_out/pytest_r2d.log 2>&1
tail -3 gpurun_out/pytest_r2d.log
HPRI_TIMELINE=gpurun_out/timeline_r2d.json python tools/ab_step.py 5 10 > gpurun_out/ab_r2d.jsonl 2> gpurun_out/ab_r2d.err
cat gpurun_out/ab_r2d.jsonl; tail -3 gpurun_out/ab_r2d.err
