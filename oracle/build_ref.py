"""Materialise `oracle/_ref/`: the UNMODIFIED reference package, importable on the GPU box.

    python oracle/build_ref.py          # needs /root/reference (the build container)

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/ve_oracle.py's header): `oracle/_ref/` is the
reference's own CPU implementation, used as the CPU baseline of `bench.py` (`cpu_baseline.kind
== "reference"`, `--impl reference`) and to validate the restatement.  Nothing under
`sorobn_b200/` imports it.

The reference (MaxHalford/sorobn) is pure Python over pandas: there is nothing to compile, the
"build" is a file copy of its four library modules from where they lie under /root/reference
into `oracle/_ref/sorobn/` (git-ignored: the sources never enter this repo's history; not
gpurun-ignored: the directory travels to the GPU box like a built .so would).  `gui.py`
(streamlit) and the reference's tests are left out.

`vose` (a Cython alias sampler the reference imports at module level, pinned `vose>=0.0.1` in
the reference's pyproject.toml) is absent from this image and has no wheel in the offline
wheelhouse.  `oracle/_ref/vose.py` is written here as a stand-in with the same two-call
interface (`Sampler(weights, seed)`, `.sample()`): inverse-CDF sampling with numpy.  Exact
inference never touches it; the sampling algorithms draw the same distribution through it (not
the same stream -- the reference's own stream is not reproducible across vose builds either).
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/sorobn"
DST = os.path.join(HERE, "_ref")
MODULES = ("__init__.py", "bayes_net.py", "examples.py", "structure.py")

VOSE_STUB = '''"""Stand-in for the `vose` package (written by oracle/build_ref.py, not part of the reference)."""
import numpy as np


class Sampler:
    def __init__(self, weights, seed=None):
        w = np.asarray(weights, dtype=float)
        self._cdf = np.cumsum(w / w.sum())
        self._rng = np.random.default_rng(seed)

    def sample(self, k=None):
        if k is None:
            return int(min(np.searchsorted(self._cdf, self._rng.random(), side="right"), len(self._cdf) - 1))
        return np.minimum(np.searchsorted(self._cdf, self._rng.random(k), side="right"), len(self._cdf) - 1)
'''


def available() -> bool:
    return os.path.exists(os.path.join(DST, "sorobn", "bayes_net.py"))


def build(force: bool = False) -> str | None:
    """Copy the reference package.  Returns the directory, or None when /root/reference is
    absent (the GPU box: the directory built in the container is used as it arrived)."""
    if not os.path.isdir(REF_SRC):
        return DST if available() else None
    pkg = os.path.join(DST, "sorobn")
    os.makedirs(pkg, exist_ok=True)
    for m in MODULES:
        src, dst = os.path.join(REF_SRC, m), os.path.join(pkg, m)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
    with open(os.path.join(DST, "vose.py"), "w") as f:
        f.write(VOSE_STUB)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Copy of /root/reference/sorobn made by oracle/build_ref.py (git-ignored). vose.py is a stand-in.\n")
    return DST


def import_reference():
    """Import the reference package from oracle/_ref (raises ImportError when it was not built)."""
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import sorobn  # noqa: E402

    if os.path.dirname(os.path.abspath(sorobn.__file__)) != os.path.join(DST, "sorobn"):
        raise ImportError(f"`sorobn` resolved to {sorobn.__file__}, not to oracle/_ref")
    return sorobn


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
