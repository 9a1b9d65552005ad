set -x
rm -f gpurun_out/parity_records.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_r2r.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2r.log
tail -n 4 gpurun_out/pytest_r2r.log
grep bilinear gpurun_out/parity_records.jsonl
