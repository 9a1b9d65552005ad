# round-2 (p): what the driver runs at round end -- the GPU suite, smoke(), the default bench line, the reference arm
set -x
rm -f gpurun_out/parity_records.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_r2p.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2p.log
tail -n 4 gpurun_out/pytest_r2p.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2p.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2p.log; tail -n 2 gpurun_out/smoke_r2p.log
timeout 600 python bench.py > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r2p.json 2> gpurun_out/bench_ref_r2p.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_r2p.json | cut -c1-600
tail -n 3 gpurun_out/bench_r2p.err gpurun_out/bench_ref_r2p.err
