"""Builds dot_b200/libdotgpu.so (sm_100a only) with nvcc, in-tree.  `python -m dot_b200.build [--force]`."""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdotgpu.so")
SOURCES = ["energy_kernels.cu", "chol_numeric.cu", "chol_solve.cu", "linalg.cu", "stepper.cu", "capi.cu", "mesh_host.cpp", "chol_symbolic.cpp",
           "anim_host.cpp", "comm.cpp", "partition_host.cpp", "peer_reduce.cu"]
METIS_LIB = os.path.join(HERE, "libdotmetis.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp,-O3",
              "-Xptxas", "-v"]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "dotgpu.h"))
    cc = nvcc()

    def compile_one(src: str) -> str:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        if not force and _newer(o, [s] + headers):
            return o
        cmd = [cc] + NVCC_FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(o + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
        return o

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or not _newer(LIB, objs):
        cmd = [cc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fopenmp", "-lgomp", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


def build_metis(force: bool = False) -> str | None:
    """libdotmetis.so = METIS 5.1.0 exactly as the reference vendors it (SuiteSparse/metis-5.1.0: IDXTYPEWIDTH 64, REALTYPEWIDTH 32),
    compiled from the sources WHERE THEY LIE under $DOT_REFERENCE (default /root/reference) - nothing is copied into the repo.
    dotgpu_partition dlopen()s it (bit-exact labels need the same METIS).  Returns None when the sources are absent
    (GPU box: the prebuilt .so travels with the snapshot)."""
    ref = os.environ.get("DOT_REFERENCE", "/root/reference")
    M = os.path.join(ref, "SuiteSparse", "metis-5.1.0")
    if not os.path.isdir(os.path.join(M, "libmetis")):
        return METIS_LIB if os.path.exists(METIS_LIB) else None
    if os.path.exists(METIS_LIB) and not force:
        return METIS_LIB
    obj = os.path.join(OBJ, "metis")
    os.makedirs(obj, exist_ok=True)
    cc = shutil.which("gcc") or "gcc"
    flags = ["-O2", "-w", "-fPIC", "-DLINUX", "-D_FILE_OFFSET_BITS=64", "-DNDEBUG", "-DNDEBUG2", "-DHAVE_EXECINFO_H", "-DHAVE_GETLINE",
             "-std=c99", "-D_GNU_SOURCE", "-I" + os.path.join(M, "GKlib"), "-I" + os.path.join(M, "include"), "-I" + os.path.join(M, "libmetis")]
    jobs = []
    for tag, d in (("gk", "GKlib"), ("lm", "libmetis")):   # both directories have util.c / graph.c / ...: prefix the objects
        for f in sorted(os.listdir(os.path.join(M, d))):
            if f.endswith(".c"):
                jobs.append((os.path.join(M, d, f), os.path.join(obj, "%s_%s.o" % (tag, f[:-2]))))

    def cc_one(j):
        r = subprocess.run([cc] + flags + ["-c", j[0], "-o", j[1]], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed on %s:\n%s" % (j[0], r.stderr))
        return j[1]

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc_one, jobs))
    r = subprocess.run([cc, "-shared", "-o", METIS_LIB] + objs + ["-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("linking libdotmetis.so failed:\n" + r.stderr)
    return METIS_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_metis(force="--force" in sys.argv))
