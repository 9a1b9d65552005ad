# round-2 (b): the whole GPU test-suite on the refactored engine, then the default bench with a per-kernel breakdown
set -x
rm -f gpurun_out/parity_records.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/pytest_r2b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
tail -40 gpurun_out/pytest_r2b.log
timeout 600 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/bd_r2b.json > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
tail -3 gpurun_out/bench_r2b.err
cat gpurun_out/bench_r2b.json
