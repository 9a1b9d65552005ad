# round-2 (r): the GPU suite on the rebuilt library, the default bench line, and the cost of the deterministic mode
set -x
rm -f gpurun_out/parity_records.jsonl
timeout 400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_r2r.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2r.log
tail -n 4 gpurun_out/pytest_r2r.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2r.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2r.log; tail -n 2 gpurun_out/smoke_r2r.log
timeout 200 python bench.py --no-extras > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; echo "bench rc=$?"
HPRI_DETERMINISTIC=1 timeout 120 python bench.py --no-extras --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/bench_det_r2r.json 2> gpurun_out/bench_det_r2r.err; echo "bench det rc=$?"
HPRI_DETERMINISTIC=fwd timeout 120 python bench.py --no-extras --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/bench_detfwd_r2r.json 2> gpurun_out/bench_detfwd_r2r.err; echo "bench det fwd rc=$?"
for f in bench_r2r bench_det_r2r bench_detfwd_r2r; do python - "$f" <<'PY'
import json, sys
l = [x for x in open(f"gpurun_out/{sys.argv[1]}.json").read().split("\n") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print(sys.argv[1], d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), d.get("api_resident", {}).get("value"), d["roofline"]["frac"])
PY
done
tail -n 3 gpurun_out/bench_r2r.err gpurun_out/bench_det_r2r.err
