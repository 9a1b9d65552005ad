# round-2 (e), two GPUs: NCCL tests (data-parallel exactness, pixel-parallel SpectralUNET), then N=2 benches
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_dp_r2e.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_dp_r2e.log
tail -5 gpurun_out/pytest_dp_r2e.log
$TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2_r2e.json 2> gpurun_out/bench_n2_r2e.err
NCCL_MAX_CTAS=4 HPRI_SM_RESERVE=4 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2_r2e_cta4.json 2>> gpurun_out/bench_n2_r2e.err
NCCL_MAX_CTAS=2 HPRI_SM_RESERVE=2 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2_r2e_cta2.json 2>> gpurun_out/bench_n2_r2e.err
NCCL_MAX_CTAS=4 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2_r2e_cta4_nores.json 2>> gpurun_out/bench_n2_r2e.err
$TR --master-port 29515 bench.py --gpus 2 --model SpectralUNET --shard pixel --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_spectral_pixel2_r2e.json 2> gpurun_out/bench_spectral_pixel2_r2e.err
python bench.py --model SpectralUNET --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bd_spectral_r2e.json > gpurun_out/bench_spectral_n1_r2e.json 2> gpurun_out/bench_spectral_n1_r2e.err
tail -3 gpurun_out/*.err
