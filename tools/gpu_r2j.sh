# round-2 (j), eight GPUs: headline config at N=8 (with e2e), what the data-parallel step pays (tools/dp_ab.py), and
# BASELINE.json configs[4] (large per-GPU batch, all-reduce hidden under backward)
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r2j.json 2> gpurun_out/bench_n8_r2j.err
timeout 150 $TR --master-port 29522 tools/dp_ab.py 3 10 2 > gpurun_out/dp_ab_n8_b2_r2j.jsonl 2> gpurun_out/dp_ab_n8_b2_r2j.err
timeout 200 $TR --master-port 29523 bench.py --gpus 8 --batch 16 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n8_b16_r2j.json 2> gpurun_out/bench_n8_b16_r2j.err
timeout 200 $TR --master-port 29524 tools/dp_ab.py 3 5 16 > gpurun_out/dp_ab_n8_b16_r2j.jsonl 2> gpurun_out/dp_ab_n8_b16_r2j.err
for f in gpurun_out/*_r2j.err; do echo == $f; grep -v "OMP_NUM\|^\*\*\*\|^$" $f | tail -n 6; done
cat gpurun_out/dp_ab_n8_b2_r2j.jsonl gpurun_out/dp_ab_n8_b16_r2j.jsonl
