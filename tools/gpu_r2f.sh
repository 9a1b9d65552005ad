# round-2 (f), two GPUs: NCCL tests (data-parallel exactness, pixel-parallel SpectralUNET), then N=2 benches
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 420 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_dp_r2f.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_dp_r2f.log
tail -15 gpurun_out/pytest_dp_r2f.log
timeout 150 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2_r2f.json 2> gpurun_out/bench_n2_r2f.err
timeout 200 $TR --master-port 29515 bench.py --gpus 2 --model SpectralUNET --shard pixel --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_spectral_pixel2_r2f.json 2> gpurun_out/bench_spectral_pixel2_r2f.err
timeout 200 python bench.py --model SpectralUNET --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bd_spectral_r2f.json > gpurun_out/bench_spectral_n1_r2f.json 2> gpurun_out/bench_spectral_n1_r2f.err
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_n1_r2f.json 2> gpurun_out/bench_n1_r2f.err
for f in gpurun_out/*_r2f.err; do echo == $f; grep -v "OMP_NUM\|^\*\*\*\|^$" $f | tail -4; done
