"""In-tree build of the native libraries (called by ``__graft_entry__.build()``).

* ``libtfcfft.so``      -- the product: sm_100a kernels + C ABI (``include/tfcfft.h``).
* ``libtfcfft_emu.so``  -- test infrastructure: the same templates executed serially on the CPU.

Explicit ``nvcc`` with ``-gencode arch=compute_100a,code=sm_100a -lineinfo``; nvcc cross-compiles
without a GPU.  The built ``.so`` files are git-ignored but travel to the GPU box with the snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(CSRC)
ROOT = os.path.dirname(PKG)

COMMON = ["-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]

TARGETS = {
    "libtfcfft.so": dict(src=["tfcfft_api.cu"], opt=["-O3"]),
    "libtfcfft_emu.so": dict(src=["emu.cu"], opt=["-O2"]),
}


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(ROOT, "include", "tfcfft.h"))
    return files


def _stale(out: str) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force: bool = False, verbose: bool = False, extra=()):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    procs = []
    for name, spec in TARGETS.items():
        out = os.path.join(PKG, name)
        if not force and not _stale(out):
            continue
        tmp = out + ".tmp"
        cmd = [nvcc, *spec["opt"], *COMMON, *extra, "-o", tmp, *[os.path.join(CSRC, s) for s in spec["src"]]]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((name, out, tmp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, out, tmp, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(log)
            raise RuntimeError(f"nvcc failed for {name}")
        os.replace(tmp, out)
        if verbose:
            print(f"built {out}", flush=True)
    return [os.path.join(PKG, n) for n in TARGETS]


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
