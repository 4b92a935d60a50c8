#!/usr/bin/env python
"""Stage the UNMODIFIED reference model into oracle/_ref/ (git-ignored; it travels to the GPU box with the
gpurun snapshot like the built .so files) -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

    python oracle/build_ref.py            # no-op when /root/reference is absent (the GPU box)

The reference is pure Python: there is nothing to compile, "building" it means placing
the files its model path consists of -- model/DSTAGNN_my.py, lib/utils.py (imported at
model/DSTAGNN_my.py:6 for scaled_Laplacian / cheb_polynomial) and lib/metrics.py (imported by lib/utils.py:8) -- where bench.py's reference arm can import
them at run time.  Nothing under oracle/_ref/ is tracked by git and nothing in the product package
(dstagnn_drought_b200/) ever imports it; only `bench.py --impl reference` and its cpu_baseline leg do, to time
the reference's own make_model('cpu', ...) on the host cores.
"""
import filecmp
import os
import shutil
import sys

REF = os.environ.get("DSTAG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["model/DSTAGNN_my.py", "lib/utils.py", "lib/metrics.py"]


def stage() -> bool:
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: keeping whatever is staged in {DST}")
        return os.path.exists(os.path.join(DST, FILES[0]))
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
        init = os.path.join(os.path.dirname(dst), "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    print(f"[build_ref] staged {', '.join(FILES)} -> {DST}")
    return True


def import_reference_make_model():
    """make_model of the staged reference (None when it is not staged).  lib/utils.py:14-16 imports
    matplotlib at module top (not installed here): stubbed, as in tests/golden/make_golden.py."""
    import types
    if not os.path.exists(os.path.join(DST, FILES[0])):
        return None
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):
        sys.modules.setdefault(name, types.ModuleType(name))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in ("model", "lib") or k.startswith(("model.", "lib."))}
    sys.path.insert(0, DST)
    try:
        from model.DSTAGNN_my import make_model
        return make_model
    finally:
        sys.path.remove(DST)
        for k in [k for k in sys.modules if k in ("model", "lib") or k.startswith(("model.", "lib."))]:
            sys.modules["_dstag_ref_" + k] = sys.modules.pop(k)       # keep them alive, out of the way
        sys.modules.update(saved)


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
