"""Results of spawned worker processes travel through multiprocessing queues BY VALUE.

A torch tensor put on a queue is sent as a file-descriptor handle the receiver has to fetch from the sender; a worker
that exits right after its `put` (as these workers do) can be gone before the parent asks, and the `get` then dies with
ConnectionResetError on a loaded machine.  `to_plain` turns every tensor of a (nested) result into a numpy array before
the `put`; `from_plain` restores tensors (and their dtype) after the `get`.
"""
import numpy as np
import torch

_TAG = "__tensor__"


def to_plain(o):
    if torch.is_tensor(o):
        t = o.detach().cpu()
        if t.dtype in (torch.bfloat16, torch.float16):
            return (_TAG, str(t.dtype), t.float().numpy().copy())
        return (_TAG, str(t.dtype), t.numpy().copy())
    if isinstance(o, dict):
        return {k: to_plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(to_plain(v) for v in o)
    return o


def from_plain(o):
    if isinstance(o, tuple) and len(o) == 3 and isinstance(o[0], str) and o[0] == _TAG:
        return torch.from_numpy(np.asarray(o[2])).to(getattr(torch, o[1].split(".")[-1]))
    if isinstance(o, dict):
        return {k: from_plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(from_plain(v) for v in o)
    return o
