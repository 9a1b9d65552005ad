# round-2 (a): streaming-kernel design probe, a fresh baseline breakdown on this pod, and a --set full capture of the
# ConvTranspose forward / dgrad launches (igemm_kernel<*,*,MODE_FWD>: 4 fwd + 4 dgrad per step), which had no ncu evidence.
set -x
./tools/probes/stream_probe 64 > gpurun_out/stream_probe_c64.txt 2>&1
./tools/probes/stream_probe 128 294272 > gpurun_out/stream_probe_c128.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/bd_r2a.json > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
python tools/prof_step.py > gpurun_out/plain_r2a.log 2>&1 &&
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:igemm_kernel<.*0>' -c 8 -o gpurun_out/prof_convT_r2a -f python tools/prof_step.py > gpurun_out/ncu_r2a.log 2>&1
ls -la gpurun_out
