set -x
timeout 300 python tools/parity_noise.py 8 > gpurun_out/parity_noise_r2.jsonl 2> gpurun_out/parity_noise_r2.err
cat gpurun_out/parity_noise_r2.jsonl; tail -n 3 gpurun_out/parity_noise_r2.err
rm -f gpurun_out/parity_records.jsonl
timeout 600 python -m pytest tests/test_zz_full_size_gpu.py tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k "full_size or deterministic or full_width" > gpurun_out/pytest_r2l.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2l.log
tail -n 5 gpurun_out/pytest_r2l.log
