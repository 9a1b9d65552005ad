"""Aggregate host->device bandwidth of N ranks copying one batch-2 fp16 cube (565 MB) each from pinned memory at the
same time, with the pinned buffer allocated (a) wherever the process happens to run and (b) after binding the process
to the CPUs of the GPU's own NUMA node (/sys/bus/pci/devices/<bus id>/numa_node).  Separates "the platform's host
memory / root complexes" from "the prefetcher" in the N=8 end-to-end number.
    torchrun --nproc-per-node N tools/h2d_probe.py"""
import json
import os

import torch
import torch.distributed as dist


def numa_cpus(dev_index):
    try:
        bus = torch.cuda.get_device_properties(dev_index).pci_bus_id
        dom = torch.cuda.get_device_properties(dev_index).pci_domain_id
        devid = torch.cuda.get_device_properties(dev_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None, node
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        return cpus, node
    except Exception as e:                      # noqa: BLE001
        return None, repr(e)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 2 * 238 * 608 * 968 * 2
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    cpus, node = numa_cpus(local)
    all_cpus = sorted(os.sched_getaffinity(0))
    out = {}
    for mode in ("default", "numa_bound"):
        if mode == "numa_bound":
            if not cpus:
                continue
            os.sched_setaffinity(0, [c for c in cpus if c in all_cpus] or all_cpus)
        src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        src.fill_(1)
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = torch.tensor([20 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9], device=dev)
        allg = [torch.zeros_like(gbs) for _ in range(world)]
        if world > 1:
            dist.all_gather(allg, gbs)
        else:
            allg = [gbs]
        out[mode] = [round(g.item(), 1) for g in allg]
        del src
        os.sched_setaffinity(0, all_cpus)
    nodes = [None] * world
    if world > 1:
        dist.all_gather_object(nodes, (local, node, len(cpus) if cpus else 0))
    else:
        nodes = [(local, node, len(cpus) if cpus else 0)]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes, "host_cpus": len(all_cpus), "gpu_numa_nodes": nodes,
                          "gb_per_s_per_rank": out, "aggregate_gb_per_s": {k: round(sum(v), 1) for k, v in out.items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
