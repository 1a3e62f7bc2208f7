"""Recipe for `oracle/_ref/`: stages the reference's OWN scatter-GEMM so that tests and the reference-equivalent GPU arm
of bench.py call the real thing instead of a restatement. TEST INFRASTRUCTURE.

The reference is Python, so "building" it means making the one module of the path that imports with what this image
has (torch + triton only) loadable on the GPU box, where /root/reference does not exist:

    /root/reference/RegionE/FluxKontext/fused_kernels.py  ->  oracle/_ref/fused_kernels.py     (byte-identical)

(all ten copies of that file in the reference are identical, md5 51f58d7e...). `oracle/_ref/` is
git-ignored - reference sources never enter the history - but travels with the snapshot to the GPU box like a built
.so. Run by `__graft_entry__.build()` whenever /root/reference is present; `python -m oracle.build_ref` by hand.
Nothing else of the reference can be staged this way: `inplace.py` subclasses diffusers classes (not installed) and
`utils.py` needs a stub of diffusers to import (oracle/make_golden.py runs it here and commits its outputs instead).
"""
from __future__ import annotations

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("RGE_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF_ROOT, "RegionE", "FluxKontext", "fused_kernels.py")
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "fused_kernels.py")


def stage(verbose: bool = True) -> str | None:
    """Copies the module if the reference tree is present; returns the staged path (or None)."""
    if not os.path.exists(SRC):
        if verbose:
            print(f"oracle/_ref: {SRC} not present (GPU box?): using what is already staged")
        return DST if os.path.exists(DST) else None
    os.makedirs(DST_DIR, exist_ok=True)
    shutil.copyfile(SRC, DST)
    digest = hashlib.md5(open(DST, "rb").read()).hexdigest()
    with open(os.path.join(DST_DIR, "MANIFEST.txt"), "w") as f:
        f.write(f"fused_kernels.py  md5 {digest}  from RegionE/FluxKontext/fused_kernels.py (unmodified)\n")
    if verbose:
        print(f"oracle/_ref: staged {DST} (md5 {digest})")
    return DST


def load_partially_linear():
    """The reference's `_partially_linear(inputs, weight, bias, index, outputs)` (fused_kernels.py:81-101), or None
    when it was never staged. Importing it needs triton; launching it needs a GPU."""
    if not os.path.exists(DST):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("regione_ref_fused_kernels", DST)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod._partially_linear


if __name__ == "__main__":
    stage()
