"""TEST INFRASTRUCTURE ONLY: build libslmholo_emu.so, the host emulation of the kernel sources.

g++ compiles the SAME .cu/.cuh files the product builds with nvcc, with -DSLM_EMULATE mapping the
CUDA constructs onto tests/emu/emu_runtime.h (one fibre per CUDA thread).  The library is used by
the CPU test-suite to check kernel logic without a GPU; the Python package never loads it.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "spatial_light_modulator_module_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libslmholo_emu.so")


def line_lengths():
    text = open(os.path.join(CSRC, "line_list.h")).read()
    return [int(x) for x in re.findall(r"X\((\d+)\)", text.split("#define SLM_LINE_LENGTHS(X)")[1])]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "emu_runtime.h")]


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force=False) -> str:
    if up_to_date() and not force:
        return LIB
    os.makedirs(OUT, exist_ok=True)
    base = ["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-DSLM_EMULATE", "-x", "c++",
            "-I", HERE, "-I", CSRC, "-Wno-unused-function", "-Wno-psabi"]
    jobs = [(os.path.join(CSRC, "engine.cu"), os.path.join(OUT, "engine.o"), []),
            (os.path.join(CSRC, "registry.cu"), os.path.join(OUT, "registry.o"), [])]
    for n in line_lengths():
        for p in (0, 1):
            jobs.append((os.path.join(CSRC, "line_inst.cu"), os.path.join(OUT, f"line_{n}_{p}.o"),
                         [f"-DSLM_LINE_L={n}", f"-DSLM_LINE_PREC={p}"]))

    def run(job):
        src, obj, defs = job
        subprocess.run(base + defs + ["-c", src, "-o", obj], check=True)
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(run, jobs))
    subprocess.run(["g++", "-shared", "-o", LIB] + objs, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
