# round-2 (m), eight GPUs: NCCL CTA budget at N=8 (NVLS needs fewer CTAs than a ring), with and without reserved SMs
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for c in 8 12 16; do
  NCCL_MAX_CTAS=$c timeout 120 $TR --master-port $((29540 + c)) tools/dp_ab.py 3 10 2 > gpurun_out/dp_ab3_n8_cta${c}_r2m.jsonl 2> gpurun_out/dp_ab3_n8_cta${c}_r2m.err
  grep "^{" gpurun_out/dp_ab3_n8_cta${c}_r2m.jsonl | sed "s/^{/{\"nccl_max_ctas\": $c, /"
done
for f in gpurun_out/*_r2m.err; do grep -v "OMP_NUM\|^\*\*\*\|^$" $f | tail -n 3; done
