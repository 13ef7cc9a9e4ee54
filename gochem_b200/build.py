"""In-tree build of libgochem_b200.so (sm_100a only) with nvcc.  Run by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libgochem_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = (["-DGCB_TRACE"] if os.environ.get("GCB_TRACE") else []) + \
    (["-DGCB_CONSUMER_WARPS=" + os.environ["GCB_CONSUMER_WARPS"]] if os.environ.get("GCB_CONSUMER_WARPS") else []) + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr"]
SOURCES = ["api.cu", "layout.cu", "super_generic.cu", "super_cluster.cu", "synth.cu", "comm.cu", "allpairs.cu", "moments.cu", "rdf.cu", "super_small.cu", "transform.cu", "lovo_ctl.cu"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "gochem_b200.h"))
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC, *ARCH, *FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl"]
        subprocess.check_call(cmd)
    return LIB


HOST_LIB = os.path.join(LIB_DIR, "libgochem_host.so")
CXX = os.environ.get("CXX", "g++")


def build_host(force: bool = False) -> str:
    """The C++ host-side mirror of the reference's Go API (gochem_b200/host), linked against the C ABI."""
    build()
    host = os.path.join(HERE, "host")
    srcs = [os.path.join(host, f) for f in ("gochem.cpp", "capi.cpp", "stf.cpp")]
    deps = srcs + [os.path.join(host, "gochem.hpp"), os.path.join(HERE, "..", "include", "gochem_host.h"),
                   os.path.join(HERE, "..", "include", "gochem_b200.h"), os.path.join(CSRC, "kabsch3.cuh")]   # capi.cpp compiles the solver for the host
    if force or _stale(HOST_LIB, deps):
        cmd = [CXX, "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-Wno-unknown-pragmas", "-o", HOST_LIB, *srcs,
               "-L" + LIB_DIR, "-lgochem_b200", "-Wl,-rpath,$ORIGIN", "-lpthread", "-lz", "-ldl"]
        subprocess.check_call(cmd)
    return HOST_LIB


if __name__ == "__main__":
    build_host()
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
