set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29561 bench.py --gpus 2 --model SpectralUNET --shard pixel --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_spectral_pixel2_r2v.json 2> gpurun_out/bench_spectral_pixel2_r2v.err
timeout 200 python bench.py --model SpectralUNET --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_spectral_n1_r2v.json 2> gpurun_out/bench_spectral_n1_r2v.err
timeout 150 $TR --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2_r2v.json 2> gpurun_out/bench_n2_r2v.err
for f in gpurun_out/*_r2v.err; do grep -v "OMP_NUM\|^\*\*\*\|^$" $f | tail -n 3; done
