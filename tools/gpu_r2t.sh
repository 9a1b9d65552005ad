set -x
timeout 400 python -m pytest tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k "nobn or without_batchnorm or SpectralUNET" > gpurun_out/pytest_r2t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2t.log
tail -n 25 gpurun_out/pytest_r2t.log
