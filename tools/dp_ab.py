"""What the data-parallel step pays over the single-GPU step, measured under torchrun on N GPUs with the configurations
interleaved in one process per rank (max over ranks per configuration):
  single        no hook: one gradient unpack launch at the end of backward (what N=1 runs)
  buckets_noop  the data-parallel code path (per-bucket unpack on the side stream, hook called) with a hook that does
                nothing -- isolates the cost of bucketing itself
  dp            the real bucketed NCCL all-reduce
  dp_reserve8   the same with 8 SMs kept out of the persistent tensor-core grids (room for NCCL's CTAs)
  dp_at_end     no overlap at all: one all-reduce of the whole gradient arena after backward
  dp_two        two all-reduces: everything finished by the middle of backward (decoder + deepest encoder block, 87 % of
                the bytes) in one call issued there, the rest at the end
    torchrun --nproc-per-node N tools/dp_ab.py [rounds] [steps] [batch]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperpri_b200 import ops, parallel                                # noqa: E402
from hyperpri_b200.src.Experiments.models import CubeNET              # noqa: E402

H, W, BANDS = 608, 968, 238


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(rank)
    net = CubeNET(BANDS, 1, first_depth=64, bilinear=False).to(dev).train()
    x = torch.rand((batch, 1, BANDS, H, W), device=dev)
    mask = (torch.rand((batch, 1, H, W), device=dev) > 0.95).float()
    eng = net._get_engine(dev)
    red = parallel.BucketedAllReduce(None, eng)
    nbytes = [0]

    def noop(flat):
        nbytes[0] += flat.numel() * 4

    mode = {"end": False, "two": False}
    pend = []

    def two_hook(flat):
        """Coalesce the per-bucket calls into two all-reduces (buckets are contiguous in the arena, in completion order)."""
        a = (flat.data_ptr() - eng.arena.data_ptr()) // 4
        b = a + flat.numel()
        mid = eng.bucket_bounds[4][1]                      # end of the deepest encoder block's bucket
        if b == mid:
            pend.append(dist.all_reduce(eng.arena[:mid], async_op=True))
        elif b == eng.arena.numel():
            pend.append(dist.all_reduce(eng.arena[mid:], async_op=True))

    def step():
        eng.invalidate_packed()
        logits = eng.forward(x, True)
        _, dlogit, _ = eng.loss_and_dlogit(logits, mask, grad_scale=1.0 / world)
        eng.set_next_input(x)
        eng.backward(dlogit, prescaled=True)
        if mode["end"]:
            dist.all_reduce(eng.arena)
        for w in pend:
            w.wait()
        pend.clear()
        red.finish()

    def cfg(hook, reserve=0, end=False):
        def apply():
            eng.bucket_hook = hook
            ops.set_sm_reserve(reserve)
            mode["end"] = end
        return apply
    configs = {"single": cfg(None), "buckets_noop": cfg(noop)}
    if world > 1:
        configs["dp"] = cfg(red.hook)
        configs["dp_reserve8"] = cfg(red.hook, 8)
        configs["dp_at_end"] = cfg(None, 0, end=True)
        configs["dp_two"] = cfg(two_hook)
    res = {k: [] for k in configs}
    for _ in range(rounds):
        for name, apply in configs.items():
            apply()
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name].append(t.item())
    if rank == 0:
        for name, v in res.items():
            v = sorted(v)
            print(json.dumps({"config": name, "n_gpus": world, "batch_per_gpu": batch, "median_ms": v[len(v) // 2], "min_ms": v[0],
                              "max_ms": v[-1], "images_per_s_median": batch * world * 1e3 / v[len(v) // 2]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
