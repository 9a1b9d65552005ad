# round-2 (k), eight GPUs: where to put the gradient all-reduce (overlapped buckets / two calls / one call at the end),
# under NCCL's default algorithm choice and with NVLS forced; NCCL_DEBUG=INFO of rank 0 kept as evidence of the algorithm
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 150 $TR --master-port 29531 tools/dp_ab.py 3 10 2 > gpurun_out/dp_ab2_n8_b2_r2k.jsonl 2> gpurun_out/dp_ab2_n8_b2_r2k.err
NCCL_ALGO=NVLS timeout 150 $TR --master-port 29532 tools/dp_ab.py 3 10 2 > gpurun_out/dp_ab2_n8_b2_nvls_r2k.jsonl 2> gpurun_out/dp_ab2_n8_b2_nvls_r2k.err
timeout 150 $TR --master-port 29533 tools/dp_ab.py 2 5 16 > gpurun_out/dp_ab2_n8_b16_r2k.jsonl 2> gpurun_out/dp_ab2_n8_b16_r2k.err
grep -h "^{" gpurun_out/dp_ab2_*r2k.jsonl
grep -i -m 12 "nvls\|algo\|channels" gpurun_out/dp_ab2_n8_b2_r2k.jsonl gpurun_out/dp_ab2_n8_b2_r2k.err | cut -c1-220
grep -v "OMP_NUM\|^\*\*\*\|^$\|NCCL INFO" gpurun_out/dp_ab2_n8_b2_nvls_r2k.err | tail -n 5
# keep the NCCL log small
grep -i "nvls\|Using network\|Channel 00\|algo\|nranks\|comm 0x" gpurun_out/dp_ab2_n8_b2_r2k.jsonl gpurun_out/dp_ab2_n8_b2_r2k.err | head -n 80 > gpurun_out/nccl_info_n8_r2k.txt
grep "^{" gpurun_out/dp_ab2_n8_b2_r2k.jsonl > gpurun_out/tmp && mv gpurun_out/tmp gpurun_out/dp_ab2_n8_b2_r2k.jsonl
rm -f gpurun_out/dp_ab2_n8_b2_r2k.err
