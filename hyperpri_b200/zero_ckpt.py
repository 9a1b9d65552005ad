"""Import of DeepSpeed ZeRO stage-2 checkpoint directories (reference src/PLTrainer.py:186-216).

The reference calls `deepspeed.utils.zero_to_fp32.convert_zero_checkpoint_to_fp32_state_dict` (DeepSpeed 0.14.2,
`test_models.ipynb:200`) and strips `_forward_module.m_network.` from the keys.  DeepSpeed is not installable here, so
this module restates the stage-2 part of that script's published algorithm:

  <dir>/latest                              text file naming the tag sub-directory (Lightning: "checkpoint")
  <dir>/<tag>/mp_rank_00_model_states.pt    {"module": 16-bit state dict, "buffer_names": [...],
                                             "param_shapes": [OrderedDict(name -> shape) per optimizer group],
                                             "shared_params": {alias: source}, ...}
  <dir>/<tag>/*_optim_states.pt             one per data-parallel rank: {"optimizer_state_dict": {"zero_stage": 2,
                                             "partition_count": world, "single_partition_of_fp32_groups": [flat fp32
                                             partition per group], ...}}

fp32 master weights of a group = the ranks' partitions concatenated in rank order, then cut sequentially by
`param_shapes`; the tail padding is what rounds the group up to a multiple of 2 * world.  Buffers come from "module".

PARITY UNPINNED: no DeepSpeed-written checkpoint exists in this environment; `tests/test_host_cpu.py` builds a
directory to this layout and checks the round trip, which pins the code to the layout above, not to DeepSpeed.
"""
from __future__ import annotations

import glob
import math
import os
import re
from collections import OrderedDict
from typing import Dict

import torch


def _natural(s: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def _numel(shape) -> int:
    return int(math.prod(tuple(shape))) if len(tuple(shape)) else 1


def fp32_state_dict_from_zero2(ckpt_dir: str, tag: str | None = None) -> "OrderedDict[str, torch.Tensor]":
    if tag is None:
        latest = os.path.join(ckpt_dir, "latest")
        if not os.path.isfile(latest):
            raise ValueError(f"no 'latest' file in {ckpt_dir}: pass the tag sub-directory explicitly")
        with open(latest) as f:
            tag = f.read().strip()
    d = os.path.join(ckpt_dir, tag)
    optim_files = sorted(glob.glob(os.path.join(d, "*_optim_states.pt")), key=_natural)
    model_files = sorted(glob.glob(os.path.join(d, "*_model_states.pt")), key=_natural)
    if not optim_files or not model_files:
        raise FileNotFoundError(f"{d} holds no *_optim_states.pt / *_model_states.pt files")
    flats, world = [], None
    for f in optim_files:
        osd = torch.load(f, map_location="cpu", weights_only=False)["optimizer_state_dict"]
        if int(osd["zero_stage"]) > 2:
            raise NotImplementedError("ZeRO stage 3 checkpoints are not produced by the reference (PLTrainer.py:414-421: stage 2)")
        pc = osd["partition_count"]
        pc = max(pc) if isinstance(pc, (list, tuple)) else int(pc)
        world = pc if world is None else world
        if pc != world:
            raise ValueError("ranks disagree on partition_count")
        flats.append([t.float() for t in osd["single_partition_of_fp32_groups"]])
    if world != len(optim_files):
        raise ValueError(f"expected {world} *_optim_states.pt files (partition_count), found {len(optim_files)}")
    ms = torch.load(model_files[0], map_location="cpu", weights_only=False)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k in ms.get("buffer_names", []):
        out[k] = ms["module"][k].float() if ms["module"][k].dtype.is_floating_point else ms["module"][k]
    shapes = ms["param_shapes"]
    if isinstance(shapes, dict):
        shapes = [shapes]
    if len(shapes) != len(flats[0]):
        raise ValueError("number of parameter groups differs between model and optimizer states")
    align = 2 * world
    for g, group in enumerate(shapes):
        full = torch.cat([flats[r][g] for r in range(world)], 0)
        off = 0
        for name, shape in group.items():
            n = _numel(shape)
            if off + n > full.numel():
                raise ValueError(f"group {g}: partitions hold {full.numel()} values, parameter {name} needs more")
            out[name] = full.narrow(0, off, n).view(tuple(shape)).clone()
            off += n
        if math.ceil(off / align) * align != math.ceil(full.numel() / align) * align:
            raise ValueError(f"group {g}: consumed {off} of {full.numel()} values (more than alignment padding left over)")
    for alias, src in (ms.get("shared_params") or {}).items():
        if src in out:
            out[alias] = out[src]
    return out


def consolidate_deepspeed_two(ckpt_path: str) -> Dict[str, torch.Tensor]:
    """PLTrainer.py:186-216: network state dict of a Lightning + DeepSpeed ZeRO-2 checkpoint directory (keys without
    `_forward_module.m_network.`; the unused `feat_ext` entries dropped as the reference does)."""
    sd = fp32_state_dict_from_zero2(ckpt_path)
    out = {}
    for k, v in sd.items():
        nk = k.replace("_forward_module.m_network.", "")
        if "feat_ext" in nk:
            continue
        out[nk] = v
    return out
