set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 60 python tools/h2d_probe.py > gpurun_out/h2d_n1_r2o.json 2> gpurun_out/h2d_n1_r2o.err
timeout 120 $TR --master-port 29551 tools/h2d_probe.py > gpurun_out/h2d_n8_r2o.json 2> gpurun_out/h2d_n8_r2o.err
grep "^{" gpurun_out/h2d_n1_r2o.json gpurun_out/h2d_n8_r2o.json
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/h2d_n8_r2o.err | tail -n 5
nproc; lscpu | grep -i "numa\|socket\|model name" | head -n 8
