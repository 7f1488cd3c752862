"""In-tree build of the native libraries (called by ``__graft_entry__.build()``).

* ``libtfcfft.so``      -- the product: sm_100a kernels + C ABI (``include/tfcfft.h``), built from several translation
                           units (``k_*.cu`` + ``tfcfft_api.cu``) compiled in parallel and linked with ``nvcc -shared``.
* ``libtfcfft_emu.so``  -- test infrastructure: the same templates executed serially on the CPU.

Explicit ``nvcc`` with ``-gencode arch=compute_100a,code=sm_100a -lineinfo``; nvcc cross-compiles
without a GPU.  The built ``.so`` / ``.o`` files are git-ignored but travel to the GPU box with the snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(CSRC)
ROOT = os.path.dirname(PKG)
OBJ = os.path.join(CSRC, "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", *ARCH, "-lineinfo"]

TARGETS = {
    # (source, extra defines): the FFT kernel families are compiled once per element type (-DTFC_DT=0..3)
    "libtfcfft.so": dict(src=[("tfcfft_api.cu", None), ("k_misc.cu", None)] +
                             [(f, dt) for f in ("k_line.cu", "k_sub.cu", "k_resident.cu", "k_split.cu") for dt in range(4)],
                         opt=["-O3"]),
    "libtfcfft_emu.so": dict(src=[("emu.cu", None)] + [("emu_fft.cu", dt) for dt in range(4)], opt=["-O2"]),
}


def _headers():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    files.append(os.path.join(ROOT, "include", "tfcfft.h"))
    return files


def _newer(path: str, deps) -> bool:
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(f) > t for f in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return p.returncode, p.stdout


def build(force: bool = False, verbose: bool = False, extra=(), only=None):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ, exist_ok=True)
    headers = _headers()
    jobs = []  # (obj, cmd)
    links = []
    for name, spec in TARGETS.items():
        if only and name not in only:
            continue
        out = os.path.join(PKG, name)
        objs = []
        for src, dt in spec["src"]:
            tag = "" if dt is None else f"_dt{dt}"
            obj = os.path.join(OBJ, name.replace(".so", "") + "__" + src.replace(".cu", tag + ".o"))
            objs.append(obj)
            defs = [] if dt is None else [f"-DTFC_DT={dt}"]
            if force or _newer(obj, [os.path.join(CSRC, src), *headers]):
                jobs.append((obj, [nvcc, *spec["opt"], *COMMON, *defs, *extra, "-c", "-o", obj, os.path.join(CSRC, src)]))
        links.append((name, out, objs))
    failed = []
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        for (obj, cmd), (rc, log) in zip(jobs, ex.map(lambda j: _run(j[1], verbose), jobs)):
            if verbose and log.strip():
                print(log, flush=True)
            if rc != 0:
                if os.path.exists(obj):
                    os.remove(obj)
                sys.stderr.write(log)
                failed.append(obj)
    if failed:
        raise RuntimeError(f"nvcc failed for {', '.join(os.path.basename(f) for f in failed)}")
    for name, out, objs in links:
        if not force and not _newer(out, objs):
            continue
        tmp = out + ".tmp"
        rc, log = _run([nvcc, "-shared", *ARCH, "-o", tmp, *objs], verbose)
        if rc != 0:
            sys.stderr.write(log)
            raise RuntimeError(f"link failed for {name}")
        os.replace(tmp, out)
        if verbose:
            print(f"built {out}", flush=True)
    return [os.path.join(PKG, n) for n in TARGETS]


def build_variant(tag: str, defines, units=("k_line.cu",), verbose: bool = False):
    """A/B build: ``libtfcfft_<tag>.so`` = the product library with ``units`` recompiled under extra ``defines``
    (e.g. ``["-DTFCFFT_LINE_BIN_EVALS=2"]``); every other object is shared with the main build.  Select it at run
    time with ``TFCFFT_LIB=<path>``.  Not part of ``build()``: variants are experiments, not the product."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    build(only=["libtfcfft.so"], verbose=verbose)
    spec = TARGETS["libtfcfft.so"]
    objs, jobs = [], []
    for src, dt in spec["src"]:
        tagd = "" if dt is None else f"_dt{dt}"
        base = os.path.join(OBJ, "libtfcfft__" + src.replace(".cu", tagd + ".o"))
        if src in units:
            obj = base.replace(".o", f"__{tag}.o")
            defs = ([] if dt is None else [f"-DTFC_DT={dt}"]) + list(defines)
            if _newer(obj, [os.path.join(CSRC, src), *_headers()]):
                jobs.append((obj, [nvcc, *spec["opt"], *COMMON, *defs, "-c", "-o", obj, os.path.join(CSRC, src)]))
            objs.append(obj)
        else:
            objs.append(base)
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        for (obj, cmd), (rc, log) in zip(jobs, ex.map(lambda j: _run(j[1], verbose), jobs)):
            if rc != 0:
                sys.stderr.write(log)
                raise RuntimeError(f"nvcc failed for {obj}")
    out = os.path.join(PKG, f"libtfcfft_{tag}.so")
    rc, log = _run([nvcc, "-shared", *ARCH, "-o", out, *objs], verbose)
    if rc != 0:
        sys.stderr.write(log)
        raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=[a for a in sys.argv[1:] if a.startswith("-X")])
