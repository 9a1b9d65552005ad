"""Minimal ENVI (.hdr + .dat) reader: the `spectral` package the reference uses
(src/dataset.py:17,262-266) is not available offline.  Supports BSQ / BIL / BIP interleaves and the
common data types; returns a float32 array in ENVI's lines x samples x bands order, which is what
`spectral`'s ``open(...).load()`` yields."""
from __future__ import annotations

import numpy as np

_DTYPES = {1: np.uint8, 2: np.int16, 3: np.int32, 4: np.float32, 5: np.float64, 12: np.uint16, 13: np.uint32,
           14: np.int64, 15: np.uint64}


def read_header(hdr_path: str) -> dict:
    txt = open(hdr_path, "r", errors="replace").read()
    if not txt.lstrip().upper().startswith("ENVI"):
        raise ValueError(f"{hdr_path}: not an ENVI header")
    out, key, buf, depth = {}, None, "", 0
    for line in txt.splitlines()[1:]:
        if depth == 0:
            if "=" not in line:
                continue
            key, val = line.split("=", 1)
            key, val = key.strip().lower(), val.strip()
            if val.startswith("{") and "}" not in val:
                depth, buf = 1, val
                continue
            out[key] = val.strip("{} \t")
        else:
            buf += " " + line.strip()
            if "}" in line:
                out[key] = buf.strip("{} \t")
                depth = 0
    return out


def load(hdr_path: str, dat_path: str) -> np.ndarray:
    """lines x samples x bands float32 (memory-mapped read, one copy)."""
    h = read_header(hdr_path)
    lines, samples, bands = int(h["lines"]), int(h["samples"]), int(h["bands"])
    dt = np.dtype(_DTYPES[int(h.get("data type", 4))])
    dt = dt.newbyteorder(">" if int(h.get("byte order", 0)) == 1 else "<")
    off = int(h.get("header offset", 0))
    inter = h.get("interleave", "bsq").lower()
    mm = np.memmap(dat_path, dtype=dt, mode="r", offset=off)
    if inter == "bsq":
        arr = mm[: bands * lines * samples].reshape(bands, lines, samples).transpose(1, 2, 0)
    elif inter == "bil":
        arr = mm[: bands * lines * samples].reshape(lines, bands, samples).transpose(0, 2, 1)
    elif inter == "bip":
        arr = mm[: bands * lines * samples].reshape(lines, samples, bands)
    else:
        raise ValueError(f"unknown interleave {inter}")
    return np.asarray(arr, dtype=np.float32)


def save(hdr_path: str, dat_path: str, cube_lsb: np.ndarray, interleave: str = "bil"):
    """Write a lines x samples x bands float32 cube (used by tests to build synthetic datasets)."""
    lines, samples, bands = cube_lsb.shape
    a = np.asarray(cube_lsb, dtype="<f4")
    data = {"bsq": a.transpose(2, 0, 1), "bil": a.transpose(0, 2, 1), "bip": a}[interleave]
    np.ascontiguousarray(data).tofile(dat_path)
    with open(hdr_path, "w") as f:
        f.write("ENVI\nsamples = %d\nlines = %d\nbands = %d\nheader offset = 0\nfile type = ENVI Standard\n"
                "data type = 4\ninterleave = %s\nbyte order = 0\n" % (samples, lines, bands, interleave))
