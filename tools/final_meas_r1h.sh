set -x
python bench.py > gpurun_out/bench_default_r1h.json 2> gpurun_out/bench_default_r1h.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bench_breakdown_r1h.json > gpurun_out/bench_10step_r1h.json 2>/dev/null
python bench.py --model UNET --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_unet_r1h.json 2>/dev/null
python bench.py --model SpectralUNET --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_spectral_r1h.json 2>/dev/null
python bench.py --mode infer --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_infer_cubenet_r1h.json 2>/dev/null
python bench.py --batch 16 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_b16_r1h.json 2>/dev/null
python tools/parity_report.py --model CubeNET --n 2 --h 608 --w 968 --json gpurun_out/parity_r1h_cubenet_2x238x608x968.json > /dev/null 2>&1
python tools/parity_report.py --model UNET --n 2 --h 608 --w 968 --json gpurun_out/parity_r1h_unet_2x3x608x968.json > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r1h.log 2>&1
tail -1 gpurun_out/smoke_r1h.log
