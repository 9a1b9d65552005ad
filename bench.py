#!/usr/bin/env python
"""Headline benchmark: train images/s (238x608x968 HSI cubes, forward + loss + backward) of CubeNET-64,
batch 2 per GPU, data-parallel over N B200s (BASELINE.json configs[2]; metric quoted at 1/2/4/8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

For N > 1 launch with torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.
  value     images/s with the fp32 cube already resident in HBM (ingest, weight re-pack, forward, BCE, backward,
            gradient all-reduce are all inside the timed region; optimizer excluded, as SURVEY.md section 8d defines)
  e2e       the same through the public nn.Module API with PINNED HOST inputs: H2D of every step's cube + mask and
            a D2H read of the loss inside the timed region (double-buffered copy stream)
  roofline  the tcgen05 implicit-GEMM family (every conv / convT fwd, dgrad, wgrad launch): algorithmic FLOPs per
            step / summed CUDA-event durations of those launches, against the measured sustained bf16 peak
  cpu_baseline  the reference's own model code (oracle/_ref, staged by oracle/build_ref.py) on this box's host cores,
            same batch and frame, bounded number of steps (falls back to the oracle port when the copy is not staged)
--impl reference times that CPU path alone (the reference is pure Python/PyTorch).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic GFLOP / image: (fwd+bwd, fwd only); SpectralUNET-1650 at its 608x700 patch
GF_PER_IMG = {"CubeNET": (2910.2, 1023.8), "UNET": (2591.5, 864.5), "SpectralUNET": (77150.9, 25828.4)}
H, W, BANDS = 608, 968, 238
PATCH_W = {"CubeNET": 968, "UNET": 968, "SpectralUNET": 700}     # params_HyperPRI.py patch sizes


def peaks():
    """(sustained bf16 TF/s, burst bf16 TF/s, HBM GB/s, source)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (d.get("bf16_tflops_sustained", 1393.2), d.get("bf16_tflops", 1640.6), d.get("hbm_gbs", 6543.7),
                "measured (MEASURED_PEAKS.json)")
    return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi polled every 10 ms from before the warm-up; samples are attributed to the timed region by their
    timestamp (the region is ~0.1-0.2 s, shorter than nvidia-smi's start-up)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, f"/tmp/hpri_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "10"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[2]), float(parts[3]), float(parts[4]), parts[5:9]))
            except ValueError:
                continue
        sel, window = rows, "warm-up + timed region"
        if t_begin is not None:
            inside = [r for r in rows if t_begin <= r[0] <= t_end]
            if inside:
                sel, window = inside, "timed region"
        if sel:
            sm = sorted(r[1] for r in sel)
            reasons = set()
            for r in sel:
                for nm, v in zip(names, r[4]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[2] for r in sel), "reasons": sorted(reasons),
                   "samples": len(sel), "window": window, "power_w_max": max(r[3] for r in sel)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


METRIC = {"train": "train images/s (238x608x968 HSI, fwd+bwd)", "infer": "inference images/s (238x608x968 HSI, fwd)"}
WORKLOAD = {
    "CubeNET": "CubeNET-64 n_channels=238 (hsi 25..263), patch 608x968, batch {n} per GPU, data-parallel "
               "(BASELINE.json configs[2])",
    "UNET": "UNET RGB n_channels=3, patch 608x968, batch {n} per GPU (BASELINE.json configs[0] shape, on GPU)",
    "SpectralUNET": "SpectralUNET n_channels=238, patch 608x700, spectral_bn_size=1650, batch {n} per GPU "
                    "(BASELINE.json configs[1]; the reference's MODEL_SHARD is ZeRO-2 data parallelism)",
}


# ------------------------------------------------------------------------------------------- CPU arm
# The CPU leg times the REFERENCE's own modules (oracle/_ref, staged by oracle/build_ref.py from
# /root/reference/src/Experiments/{models,model_parts}.py; kind "reference") at the benchmark's own batch and frame:
# batch 2, full 608 rows, fp32, all host threads.  Without the staged copy it falls back to the oracle port (kind
# "port").  SpectralUNET-1650 needs ~146 GB for batch 2 at full width, so it runs on a fixed 608 x 70 strip (1/10 of
# the pixels of the 608 x 700 patch; the network is per-pixel) and the throughput is scaled by the pixel fraction.
CPU_STRIP_W = {"CubeNET": 968, "UNET": 968, "SpectralUNET": 70}


def _load_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import build_ref
        return build_ref.load_ref()
    except Exception:
        return None


def cpu_runner(model, n_img, train=True):
    """Returns (step_fn, kind, pixel_fraction, description): step_fn() runs one fwd+loss+bwd (or eval forward + loss)
    over n_img images on the host."""
    import torch
    torch.set_num_threads(os.cpu_count())
    Wp, ws = PATCH_W[model], CPU_STRIP_W[model]
    frac = ws / Wp
    bands = 3 if model == "UNET" else BANDS
    g = torch.Generator().manual_seed(0)
    x = torch.rand((n_img, bands, H, ws), generator=g)
    mask = (torch.rand((n_img, 1, H, ws), generator=g) > 0.95).float()
    what = "fwd+loss+bwd, train mode" if train else "eval-mode forward + loss"
    shape = f"{n_img} x {bands} x {H} x {ws}" + ("" if frac == 1.0 else f" ({ws}/{Wp} of the patch columns; per-pixel network, throughput scaled by the pixel fraction)")
    ref = _load_reference()
    if ref is not None:
        if model == "CubeNET":
            net, xin = ref.CubeNET(BANDS, 1, first_depth=64, bilinear=False), x[:, None]
        elif model == "UNET":
            net, xin = ref.UNet(3, 1, bilinear=False), x
        else:
            net, xin = ref.SpectralUNET(BANDS, 1, bn_feats=1650), x
        net.train(train)
        crit = torch.nn.BCEWithLogitsLoss()

        def step():
            if train:
                net.zero_grad(set_to_none=True)
                crit(net(xin), mask).backward()
            else:
                with torch.no_grad():
                    crit(net(xin), mask)
        return step, "reference", frac, f"reference modules (oracle/_ref), fp32, {shape}, {what}"
    import hyperpri_oracle as O
    if model == "CubeNET":
        schema, xin = O.unet_schema(1, 1, "cube", hsi_depth=BANDS), x[:, None]
    elif model == "UNET":
        schema, xin = O.unet_schema(3, 1, "unet"), x
    else:
        schema, xin = O.spectral_schema(BANDS, 1, 1650), x
    sd = O.synth_state_dict(schema, 0)

    def step():
        if train:
            O.forward_backward(model, xin, mask, sd, training=True)
        else:
            with torch.no_grad():
                O.bce_with_logits(O.FORWARDS[model](xin, sd, False, None), mask)
    return step, "port", frac, f"oracle port (oracle/_ref not staged), fp32, {shape}, {what}"


def cpu_time(model, n_img, train, steps, warmup, budget_s):
    """Time up to `steps` steps after `warmup` untimed ones; the step COUNT (never the frame) is cut so that the whole
    run fits `budget_s`.  Returns (images/s, seconds per step, steps timed, kind, description)."""
    step, kind, frac, desc = cpu_runner(model, n_img, train)
    t0 = time.perf_counter()
    step()                                            # first call: also the first warm-up step
    t_first = time.perf_counter() - t0
    warm = max(0, min(warmup, 1) - 1)                 # one warm-up step is what fits; it has just run
    for _ in range(warm):
        step()
    n_steps = max(1, min(steps, int((budget_s - t_first * (1 + warm)) / max(t_first, 1e-3))))
    times = []
    for _ in range(n_steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_img * frac / sec, sec, n_steps, kind, desc


def cpu_baseline(model, train, n_img):
    v, sec, n_steps, kind, desc = cpu_time(model, n_img, train, 2, 1, 30.0)
    return {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"{n_steps} timed step(s) after 1 warm-up, {sec:.2f} s/step: {desc}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    train = args.mode == "train"
    n = args.batch
    img_s, sec, n_steps, kind, desc = cpu_time(args.model, n, train, args.steps, args.warmup, 170.0)
    line = {
        "impl": "reference", "metric": METRIC[args.mode], "value": img_s, "unit": "images/s",
        "n_gpus": args.gpus, "steps": n_steps, "steps_requested": args.steps, "warmup": 1, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.model].format(n=n), "global_batch": n, "parallelism": "cpu", "mode": args.mode,
                   "reference_arm": f"{desc}; {torch.get_num_threads()} host threads; the step count (not the frame) is "
                                    f"cut to fit the time budget: {n_steps} of the requested {args.steps} steps"},
        "cpu_baseline": {"value": img_s, "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
                         "sample": f"{n_steps} timed step(s) after 1 warm-up, {sec:.2f} s/step: {desc}"},
        "e2e": {"value": img_s, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--model", default="CubeNET", choices=["CubeNET", "UNET", "SpectralUNET"])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="train: fwd+loss+bwd (headline); infer: eval-mode forward only (kfold_validate-style sweep)")
    ap.add_argument("--shard", default="batch", choices=["batch", "pixel"],
                    help="batch: data parallel, every rank its own batch (weak scaling; the headline).  pixel: SpectralUNET's "
                         "model-sharded option -- every rank a row strip of every image of ONE batch (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the short companion measurements the default N=1 run adds (other BASELINE.json configs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", default=None, help="write a per-kernel time breakdown JSON here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from hyperpri_b200 import _lib, ops, parallel
    from hyperpri_b200.src.Experiments.models import CubeNET, SpectralUNET, UNet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(3, args.warmup)

    pixel = args.shard == "pixel"
    if pixel and args.model != "SpectralUNET":
        raise SystemExit("--shard pixel is the SpectralUNET model-sharded option")
    torch.manual_seed(1234 + (0 if pixel else rank))      # pixel parallel: every rank holds the same batch
    n = args.batch
    train = args.mode == "train"
    Wp = PATCH_W[args.model]
    if args.model == "CubeNET":
        net = CubeNET(BANDS, 1, first_depth=64, bilinear=False).to(dev)
        x = torch.rand((n, 1, BANDS, H, Wp), device=dev)
    elif args.model == "UNET":
        net = UNet(3, 1, bilinear=False).to(dev)
        x = torch.rand((n, 3, H, Wp), device=dev)
    else:
        net = SpectralUNET(BANDS, 1, bn_feats=1650).to(dev)
        x = torch.rand((n, BANDS, H, Wp), device=dev)
    net.train(train)
    mask = (torch.rand((n, 1, H, Wp), device=dev) > 0.95).float()
    if pixel:
        net.enable_pixel_parallel(None)
    eng = net._get_engine(dev)
    red = parallel.attach(eng) if not pixel else parallel.BucketedAllReduce(None, None)
    gscale = red.grad_scale() if not pixel else 1.0
    jobs = 1 if pixel else world                          # batches processed per step by the whole job

    def step():
        if train:
            eng.invalidate_packed()                    # weights change every optimizer step: re-pack inside the step
        logits = eng.forward(x, train)
        if train:
            _, dlogit, _ = eng.loss_and_dlogit(logits, mask, grad_scale=gscale)
            if hasattr(eng, "set_next_input"):
                eng.set_next_input(x)                  # the next step's batch (resident in HBM) is ingested under this backward
            eng.backward(dlogit, prescaled=True)
            red.finish()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)                      # nvidia-smi start-up
    for _ in range(W_):
        step()
    sync()
    l0 = _lib.lib().hpri_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync()
    t_end = time.time()
    launches = _lib.lib().hpri_launch_count() - l0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    ms_step = ms_total / args.steps
    value = n * jobs * args.steps / (ms_total / 1e3)

    # ---------------- per-kernel durations (CUDA events around every native call; separate pass)
    # The weight-gradient side stream is switched off for this pass: with two streams sharing the SMs an event pair around
    # one kernel also counts the time it waited for the other stream's kernel, so only serialised launches give a
    # kernel's own duration.  (The timed region above runs with the overlap on.)
    prof_steps = 3
    eng.set_overlap(False)
    step()
    torch.cuda.synchronize()
    ops.PROFILE = []
    for _ in range(prof_steps):
        step()
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    eng.set_overlap(True)
    per = {}
    per_launch = []
    for name, a, b, desc in rec:
        t = a.elapsed_time(b)
        if name in ops.TENSOR_KERNELS or t > 0.05:
            per_launch.append((name, desc, t))
        d = per.setdefault(name, [0.0, 0])
        d[0] += t / prof_steps
        d[1] += 1
    tensor_ms = sum(v[0] for k, v in per.items() if k in ops.TENSOR_KERNELS)
    # median over the profiled steps (a host hiccup between two launches of one step otherwise skews the mean)
    chunk = len(rec) // prof_steps
    if chunk * prof_steps == len(rec) and chunk > 0:
        per_step = sorted(sum(a.elapsed_time(b) for name, a, b, _ in rec[i * chunk:(i + 1) * chunk] if name in ops.TENSOR_KERNELS)
                          for i in range(prof_steps))
        tensor_ms = per_step[prof_steps // 2]
    tensor_launches = sum(v[1] for k, v in per.items() if k in ops.TENSOR_KERNELS) // prof_steps
    all_ms = sum(v[0] for v in per.values())
    peak_tf, peak_burst, peak_gbs, peak_src = peaks()
    flops_step = GF_PER_IMG[args.model][0 if train else 1] * 1e9 * n / (world if pixel else 1)     # this rank's share
    achieved = flops_step / (tensor_ms / 1e3) / 1e12 if tensor_ms > 0 else 0.0
    # DRAM traffic of the same launches from the committed ncu capture (same workload only: CubeNET-64, batch 2, train)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic_r2.json")
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "traffic_r1h.json")
    if args.model == "CubeNET" and n == 2 and train and os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["tensor_family_dram_bytes_per_launch"], f"profiles/{os.path.basename(tp)} (ncu, per launch)"
    roofline = {"bound": "tensor", "kernel": "conv3x3_halo_kernel<BLOCK_N> + igemm_kernel<BLOCK_N,STAGES,MODE> (tcgen05 implicit GEMM family: every "
                          "conv3x3 / ConvTranspose / Linear fwd, dgrad and wgrad launch of the step)",
                "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved / peak_burst,
                "peak_kind": "burst bf16 cuBLAS peak: the launches are timed one at a time (serialised CUDA-event pass)",
                "frac_of_sustained_peak": achieved / peak_tf, "sustained_peak": peak_tf,
                "whole_step": {"achieved": flops_step / (ms_step / 1e3) / 1e12, "frac_of_sustained_peak": flops_step / (ms_step / 1e3) / 1e12 / peak_tf,
                               "frac_of_burst_peak": flops_step / (ms_step / 1e3) / 1e12 / peak_burst,
                               "note": "algorithmic FLOPs of the step / the timed region's ms_per_step (every kernel, overlap on)"},
                "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "flops_per_step": flops_step, "kernel_ms_per_step": tensor_ms,
                "launches_per_step": tensor_launches, "share_of_step": tensor_ms / all_ms if all_ms else None,
                "timing": "CUDA events around every launch, launches serialised (weight-gradient side stream off for this "
                          "pass); share_of_step is of the serialised sum of kernel times"}
    if args.breakdown and rank == 0:
        with open(args.breakdown, "w") as f:
            json.dump({"ms_per_step_sum_of_kernels": all_ms, "ms_per_step_wall": ms_step,
                       "kernels": {k: {"ms_per_step": v[0], "launches_per_step": v[1] / prof_steps} for k, v in
                                   sorted(per.items(), key=lambda kv: -kv[1][0])},
                       "launches_last_step": [{"kernel": k, "args": d_, "ms": t} for k, d_, t in
                                              per_launch[len(per_launch) - len(per_launch) // prof_steps:]]}, f, indent=1)

    # ---------------- end to end through the public API with host inputs
    def run_e2e(host_dtype):
        """nn.Module API, pinned host cube + mask copied H2D every step (hyperpri_b200.prefetch.DevicePrefetcher: copy
        stream, two device slots), loss read back."""
        xh = [torch.rand(x.shape).to(host_dtype).pin_memory() for _ in range(2)]
        mh = [(torch.rand(mask.shape) > 0.95).float().pin_memory() for _ in range(2)]
        from hyperpri_b200.prefetch import DevicePrefetcher
        total = W_ + args.steps

        def loader():                                   # what a DataLoader with pinned batches hands over
            for i in range(total):
                yield {"image": xh[i % 2], "mask": mh[i % 2]}

        t0 = None
        pf = DevicePrefetcher(loader(), dev)
        for i, b in enumerate(pf):                      # next step's H2D overlaps this step's compute
            if i == W_:
                sync()
                t0 = time.perf_counter()
            if train and pf.next_batch is not None:     # as the trainer loop does: its ingest runs under this backward
                net.set_next_input(pf.next_batch["image"], pf.next_ready)
            loss = api_step(b["image"], b["mask"])
            loss.item()                                 # D2H read of the step's result
        sync()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        del xh, mh
        return {"value": n * jobs * args.steps / dt.item(), "unit": "images/s",
                "h2d_bytes_per_step": x.numel() * xh_bytes[host_dtype] + mask.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": dt.item() / args.steps * 1e3}

    def api_step(xb, mb):
        """What RootLightningModel.training_step / validation_step run (hyperpri_b200/src/PLTrainer.py: _step): the
        nn.Module's bce_step = forward + BCEWithLogitsLoss + (train) backward through autograd, gradients delivered to
        the Parameters' .grad.  The weights are marked changed every step, as after an optimizer step."""
        if train:
            eng.invalidate_packed()
            net.zero_grad(set_to_none=True)
            loss, _, _ = net.bce_step(xb, mb, grad_scale=gscale)
            loss.backward()                             # the data-parallel all-reduce is finished inside
        else:
            with torch.no_grad():
                loss, _, _ = net.bce_step(xb, mb)
        return loss

    def run_api_resident():
        """The same public API with the batch already in HBM: must match `value` (same kernels, autograd on top)."""
        for _ in range(W_):
            net.set_next_input(x)
            api_step(x, mask)
        sync()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            net.set_next_input(x)
            api_step(x, mask)
        a1.record()
        sync()
        t = torch.tensor([a0.elapsed_time(a1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"value": n * jobs * args.steps / (t.item() / 1e3), "unit": "images/s", "ms_per_step": t.item() / args.steps,
                "what": "nn.Module.bce_step + loss.backward() (the trainer's step body), inputs resident in HBM"}

    api_resident = run_api_resident()
    if train:
        # the whole training iteration of the reference's loop (PLTrainer.py:79-98 + configure_optimizers :164-174):
        # zero_grad, step body, backward, Adam -- FusedAdam updates every tensor in one launch and its version bump makes
        # the next forward re-pack the fp16 operands (no invalidate_packed here)
        from hyperpri_b200.optim import FusedAdam
        opt = FusedAdam(net.parameters(), lr=1e-4, found_inf=eng.overflow)

        def full_iter():
            net.zero_grad(set_to_none=True)
            net.set_next_input(x)
            loss, _, _ = net.bce_step(x, mask, grad_scale=gscale)
            loss.backward()
            opt.step()
        for _ in range(W_):
            full_iter()
        sync()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            full_iter()
        b1.record()
        sync()
        tt = torch.tensor([b0.elapsed_time(b1)], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        api_resident["with_optimizer"] = {"value": n * jobs * args.steps / (tt.item() / 1e3), "unit": "images/s",
                                          "ms_per_step": tt.item() / args.steps,
                                          "what": "zero_grad + bce_step + backward + FusedAdam.step (one launch), inputs resident in HBM"}
        del opt
    xh_bytes = {torch.float32: 4, torch.float16: 2}
    e2e = e2e32 = None
    if not args.no_e2e:
        if args.model != "UNET":
            # the trainer's HSI loaders hand over fp16 cubes (params_HyperPRI: HyperpriDataset(host_dtype=float16), the
            # conversion is done by the loader before the copy; bit-identical network input): the default e2e format
            e2e = run_e2e(torch.float16)
            e2e["host_format"] = "fp16 cube (the trainer's default loader format: HyperpriDataset(host_dtype=float16)); bit-identical network input"
            e2e32 = run_e2e(torch.float32)          # the reference data loader's format (dataset.py:270: float32 cube)
            e2e32["host_format"] = "fp32 cube (reference dataset format); PCIe-bound: see h2d_bytes_per_step / ms_per_step"
        else:
            e2e = run_e2e(torch.float32)
            e2e["host_format"] = "fp32 RGB image (reference dataset format)"

    # ---------------- companion measurements (default N=1 run only): the other BASELINE.json configurations, a few
    # steps each, so that they are measured wherever this file is run -- configs[1] SpectralUNET-1650 training at its
    # 608 x 700 patch and configs[3] the kfold_validate-style eval-mode forward of the three models
    extras = None
    if world == 1 and args.model == "CubeNET" and train and not args.no_extras:
        red.engine, eng.bucket_hook = None, None          # drop the headline model's workspace (~8 GB) before the next ones
        net.__dict__.pop("_eng", None)
        eng.ws = None
        torch.cuda.empty_cache()
        extras = {}

        def timed(fn, steps, warm=3):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(steps):
                fn()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / steps

        def make(model_name):
            wp = PATCH_W[model_name]
            if model_name == "CubeNET":
                m_, x_ = CubeNET(BANDS, 1, first_depth=64, bilinear=False).to(dev), torch.rand((2, 1, BANDS, H, wp), device=dev)
            elif model_name == "UNET":
                m_, x_ = UNet(3, 1, bilinear=False).to(dev), torch.rand((2, 3, H, wp), device=dev)
            else:
                m_, x_ = SpectralUNET(BANDS, 1, bn_feats=1650).to(dev), torch.rand((2, BANDS, H, wp), device=dev)
            return m_, x_, (torch.rand((2, 1, H, wp), device=dev) > 0.95).float()

        for model_name in ("CubeNET", "UNET", "SpectralUNET"):
            m_, x_, k_ = make(model_name)
            e_ = m_._get_engine(dev)
            m_.eval()

            def infer():
                with torch.no_grad():
                    e_.forward(x_, False)
            ms_i = timed(infer, 10 if model_name != "SpectralUNET" else 4)
            gf = GF_PER_IMG[model_name][1] * 2
            extras[f"infer_{model_name}"] = {
                "metric": METRIC["infer"], "value": 2e3 / ms_i, "unit": "images/s", "ms_per_step": ms_i,
                "workload": WORKLOAD[model_name].format(n=2) + "; eval-mode forward (running statistics), ingest included "
                            "(BASELINE.json configs[3])",
                "whole_step_tflops": gf / ms_i, "frac_of_sustained_peak": gf / ms_i / peak_tf, "frac_of_burst_peak": gf / ms_i / peak_burst}
            if model_name == "SpectralUNET":
                m_.train()

                def train_step():
                    e_.invalidate_packed()
                    lg_ = e_.forward(x_, True)
                    _, dl_, _ = e_.loss_and_dlogit(lg_, k_)
                    e_.backward(dl_, prescaled=True)
                ms_t = timed(train_step, 3)
                gft = GF_PER_IMG[model_name][0] * 2
                extras["train_SpectralUNET"] = {
                    "metric": METRIC["train"], "value": 2e3 / ms_t, "unit": "images/s", "ms_per_step": ms_t,
                    "workload": WORKLOAD[model_name].format(n=2), "whole_step_tflops": gft / ms_t,
                    "frac_of_sustained_peak": gft / ms_t / peak_tf, "frac_of_burst_peak": gft / ms_t / peak_burst}
            m_.__dict__.pop("_eng", None)
            del e_, m_, x_, k_
            torch.cuda.empty_cache()

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline(args.model, train, n)

    if rank == 0:
        line = {
            "metric": METRIC[args.mode], "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": W_, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if pixel else "weak",
            "vs_baseline": None, "dtype": "f16 (activations, weights, loss-scaled gradients), f32 accumulate (tcgen05 kind::f16)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD[args.model].format(n=n),
                       "global_batch": n * jobs, "parallelism": (f"pixel{world} (row strips of every image, per-layer "
                                                                  "BatchNorm-statistics all-reduce)" if pixel else f"dp{world}"),
                       "mode": args.mode,
                       "l2": "inputs larger than L2 (>= 0.8 GB fp32 cube + > 3 GB activations per step); no explicit flush",
                       "timed_region": ("weight re-pack + ingest + forward + BCE + backward + grad all-reduce" if train
                                        else "ingest + eval-mode forward (running statistics)")},
            "e2e": e2e, "e2e_fp32_host": e2e32, "api_resident": api_resident, "extras": extras, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_base, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
