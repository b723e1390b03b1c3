"""Build lib/libslmholo.so for sm_100a with nvcc (cross-compiles without a GPU).

    python -m spatial_light_modulator_module_b200.build [--force] [--verbose]

One object per (line length, precision) instantiation of the pass kernels (csrc/line_inst.cu),
compiled in parallel, plus the engine / C ABI (csrc/engine.cu) and the length registry.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libslmholo.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def line_lengths():
    text = open(os.path.join(CSRC, "line_list.h")).read()
    return [int(x) for x in re.findall(r"X\((\d+)\)", text.split("#define SLM_LINE_LENGTHS(X)")[1].split("\n")[0])]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.abspath(__file__)]


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in _sources())


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None, lengths=None) -> str:
    """``defines``/``out``/``lengths`` build a tuning variant (e.g. -DSLM_TCMAX_F32=4) beside the product library."""
    global OBJ, LIB
    if out is None and up_to_date() and not force:
        return LIB
    obj_dir, lib = (OBJ, LIB) if out is None else (os.path.join(HERE, "lib", "obj_" + os.path.basename(out)), out)
    return _build(force, verbose, list(defines), obj_dir, lib, lengths)


def _build(force, verbose, defines, OBJ, LIB, lengths):
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    base = [cc, "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-I", CSRC] + ARCH + defines
    if lengths:
        base.append("-DSLM_LINE_LENGTHS(X)=" + " ".join(f"X({n})" for n in lengths))
    if verbose:
        base += ["-Xptxas", "-v"]
    jobs = [(os.path.join(CSRC, "engine.cu"), os.path.join(OBJ, "engine.o"), []),
            (os.path.join(CSRC, "registry.cu"), os.path.join(OBJ, "registry.o"), [])]
    for n in (lengths or line_lengths()):
        for p in (0, 1):
            jobs.append((os.path.join(CSRC, "line_inst.cu"), os.path.join(OBJ, f"line_{n}_{p}.o"),
                         [f"-DSLM_LINE_L={n}", f"-DSLM_LINE_PREC={p}"]))

    def run(job):
        src, obj, defs = job
        r = subprocess.run(base + defs + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {os.path.basename(obj)}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(f"==== {os.path.basename(obj)}\n{r.stderr}\n")
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(run, jobs))
    r = subprocess.run([cc, "-shared", "-o", LIB] + ARCH + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
