"""Stage the UNMODIFIED reference modules under oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            # build container only: needs /root/reference

The reference (SebastianBitsch/SRGAN-ST) is pure Python on top of torch / torchvision, which the GPU
box has too; /root/reference itself does not exist there.  This recipe copies the four files the loss
hot path and its callers import -- ``loss.py``, ``utils.py`` (the path itself), ``model.py`` (imported by
``loss.py:8``; its ``Generator`` is the producer of ``sr`` in warmup.py:86) and ``config.py`` (the
criterion registry, config.py:122-125) -- byte for byte into ``oracle/_ref/``, which is listed in
``.gitignore`` (reference sources never enter this repository's history) but not in ``.gpurunignore``
(so the staged copy travels to the GPU box with the snapshot, like the built ``.so`` files).

Who may use it (same rule as the rest of ``oracle/``): ``tests/`` (live parity against the reference
running on the same B200), ``bench.py --impl reference`` / ``cpu_baseline`` / ``reference_gpu_unfused``
(the reference's own ATen path as the timed baseline) and ``__graft_entry__``; never the product package.

``load()`` imports the staged modules under private names so they cannot shadow anything
(``loss`` / ``utils`` / ``model`` / ``config`` are common module names).
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SRST_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ["loss.py", "utils.py", "model.py", "config.py"]


def stage(verbose: bool = True) -> bool:
    """Copy the reference files; returns False (and changes nothing) when /root/reference is absent."""
    if not os.path.isdir(REF_SRC):
        if verbose:
            print(f"{REF_SRC} not present: keeping whatever is staged in {REF_DST}")
        return False
    os.makedirs(REF_DST, exist_ok=True)
    lines = []
    for f in FILES:
        src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DST, f)
        shutil.copyfile(src, dst)
        lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(REF_DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print("\n".join(lines))
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES)


_loaded = None


def load():
    """Import the staged reference; returns a namespace with .loss, .utils, .model, .config.

    The reference resolves ``from utils import ...`` / ``from model import ...`` by bare module name
    (loss.py:8-9), so ``oracle/_ref`` is put at the FRONT of sys.path for the duration of the import and
    the modules are then re-registered as ``srst_ref_*`` and removed from their bare names.
    On a box without a GPU the one shim of SURVEY.md 8c applies: ``utils.py:206,208`` hard-code ``.cuda()``.
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError(f"{REF_DST} is not staged: run `python oracle/make_ref.py` in the build container")
    import torch
    saved = {n: sys.modules.pop(n) for n in ("loss", "utils", "model", "config") if n in sys.modules}
    sys.path.insert(0, REF_DST)
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        mods = {}
        for n in ("utils", "model", "config", "loss"):
            spec = importlib.util.spec_from_file_location(n, os.path.join(REF_DST, n + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[n] = m
            spec.loader.exec_module(m)
            mods[n] = m
    finally:
        sys.path.remove(REF_DST)
        sys.dont_write_bytecode = dont
        for n in ("loss", "utils", "model", "config"):
            m = sys.modules.pop(n, None)
            if m is not None:
                sys.modules["srst_ref_" + n] = m
        sys.modules.update(saved)

    class _NS:
        pass

    ns = _NS()
    ns.loss, ns.utils, ns.model, ns.config = mods["loss"], mods["utils"], mods["model"], mods["config"]
    ns.cuda_shimmed = False
    _loaded = ns
    return ns


class on_cpu:
    """Context manager: run the reference on CPU tensors.  ``get_gaussian_kernel`` calls ``.cuda()`` on its
    taps (utils.py:206,208); inside this block ``Tensor.cuda`` is the identity so the taps stay on the
    CPU.  Used by the CPU baseline legs on the GPU box and by everything in the GPU-less container."""

    def __enter__(self):
        import torch
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._orig
        return False


if __name__ == "__main__":
    ok = stage()
    sys.exit(0 if ok or available() else 1)
