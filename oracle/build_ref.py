"""Recipe for ``oracle/_ref``: the reference's OWN model code, staged where the GPU box can run it -- TEST / BENCH
INFRASTRUCTURE ONLY.

The reference is pure Python.  Its model definitions (``src/Experiments/models.py``, ``model_parts.py``) import only
torch (models.py:10-14, model_parts.py:9-11), so they run unmodified on CPU anywhere torch does; the rest of the
reference (``src/dataset.py``, ``src/PLTrainer.py``, ``params_HyperPRI.py``) needs Lightning / torchmetrics /
DeepSpeed / spectral, none of which is installed offline, and ``pip install /root/reference`` has nothing to build
(no setup.py / pyproject).  ``/root/reference`` does not exist on the GPU box, so ``__graft_entry__.build()`` runs this
recipe in the build container: the two files are staged, byte for byte, as the package ``oracle/_ref/hyperpri_reference``
(``oracle/_ref/`` is git-ignored -- reference sources never enter the history -- but not gpurun-ignored, so the staged
copy travels).  ``bench.py --impl reference`` and the in-line ``cpu_baseline`` then time the reference modules
themselves (``kind: "reference"``); without the staged copy they fall back to the oracle port (``kind: "port"``).

Nothing under ``hyperpri_b200/`` imports this.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("HPRI_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref", "hyperpri_reference")
FILES = ("models.py", "model_parts.py")


def build_ref(force: bool = False) -> str | None:
    """Stage the reference model files; returns the package directory, or None when neither the reference nor a
    previously staged copy is present."""
    src_dir = os.path.join(REF_SRC, "src", "Experiments")
    have_src = all(os.path.exists(os.path.join(src_dir, f)) for f in FILES)
    have_dst = all(os.path.exists(os.path.join(DST, f)) for f in FILES)
    if not have_src:
        return DST if have_dst else None
    os.makedirs(DST, exist_ok=True)
    manifest = []
    for f in FILES:
        s, d = os.path.join(src_dir, f), os.path.join(DST, f)
        if force or not os.path.exists(d) or open(s, "rb").read() != open(d, "rb").read():
            shutil.copyfile(s, d)
        manifest.append(f"{hashlib.sha256(open(d, 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(DST, "__init__.py"), "w") as fh:
        fh.write('"""Unmodified copies of the reference\'s src/Experiments/{models,model_parts}.py (staged by oracle/build_ref.py)."""\n')
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    return DST


def load_ref():
    """Import the staged reference models; returns the module (``.UNet``, ``.SpectralUNET``, ``.CubeNET``) or None."""
    if not all(os.path.exists(os.path.join(DST, f)) for f in FILES):
        return None
    root = os.path.dirname(DST)
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    return importlib.import_module("hyperpri_reference.models")


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
