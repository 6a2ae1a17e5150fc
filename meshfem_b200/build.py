"""Build recipe for libmfem_b200.so (hand-written sm_100a CUDA behind the C ABI).

In-tree build with nvcc; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
LIB = os.path.join(LIBDIR, "libmfem_b200.so")
SOURCES = ["setup.cu", "assemble.cu", "solver.cu", "aux.cu", "shape.cu", "comm.cu", "capi.cu"]
HEADERS = ["core.cuh", "elem_math.cuh", "solver_multi.inl", "coarse.inl", "matfree.inl", "peer.cuh", os.path.join("..", "..", "include", "mfem_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
]


def _nccl_link_flags():
    # Prefer the NCCL bundled with torch (the one torch.distributed loads); fall back to the system one.
    cands = []
    try:
        import nvidia.nccl  # type: ignore
        cands.append(os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib"))
    except Exception:
        pass
    for d in cands:
        so = os.path.join(d, "libnccl.so.2")
        if os.path.exists(so):
            return ["-L" + d, "-l:libnccl.so.2", "-Xlinker", "-rpath=" + d]
    return ["-lnccl"]


def _stamp():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file):
        if open(stamp_file).read().strip() == stamp:
            return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"] + _nccl_link_flags()
    subprocess.check_call(cmd)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


HOST_LIB = os.path.join(LIBDIR, "libmeshfem_host.so")
REPO = os.path.dirname(ROOT)
HOST_SOURCES = [os.path.join(REPO, "src", "host", f) for f in ("MeshIO.cc", "host_capi.cc")]


def _host_stamp():
    h = hashlib.sha256()
    files = list(HOST_SOURCES)
    for d, _, fs in os.walk(os.path.join(REPO, "include")):
        files += [os.path.join(d, f) for f in fs]
    for d, _, fs in os.walk(os.path.join(REPO, "src", "bin")):
        files += [os.path.join(d, f) for f in fs]
    for f in sorted(files):
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build_host(force: bool = False) -> str:
    """Host C++ mirror of the reference's mesh layer (g++; no CUDA)."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = HOST_LIB + ".stamp"
    stamp = _host_stamp()
    if not force and os.path.exists(HOST_LIB) and os.path.exists(stamp_file):
        if open(stamp_file).read().strip() == stamp:
            return HOST_LIB
    # the host library mirrors the reference's classes; Simulator/SPSDSystem call the C ABI, so it
    # links libmfem_b200.so (build() must run first)
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-I" + os.path.join(REPO, "include")] + HOST_SOURCES + \
          ["-L" + LIBDIR, "-lmfem_b200", "-Wl,-rpath,$ORIGIN", "-o", HOST_LIB]
    subprocess.check_call(cmd)
    bindir = os.path.join(REPO, "bin")
    os.makedirs(bindir, exist_ok=True)
    for name, src in (("Simulate_cli", "src/bin/Simulate_cli.cc"),
                      ("PeriodicHomogenization_cli", "src/bin/PeriodicHomogenization_cli.cc"),
                      ("ConstStrainDisplacement_cli", "src/bin/ConstStrainDisplacement_cli.cc"),
                      ("DeformedCells_cli", "src/bin/DeformedCells_cli.cc"),
                      ("grid", "src/bin/tools/grid.cc")):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(REPO, "include"),
                               os.path.join(REPO, src), os.path.join(REPO, "src", "host", "MeshIO.cc"),
                               "-L" + LIBDIR, "-lmfem_b200", "-Wl,-rpath,$ORIGIN/../meshfem_b200/lib",
                               "-o", os.path.join(bindir, name)])
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return HOST_LIB


PY_DIR = os.path.join(REPO, "python")
PY_MODULES = (("TENSORS", "tensors"), ("SPARSE_MATRICES", "sparse_matrices"), ("HOMOGENIZATION", "periodic_homogenization"))


def build_python(force: bool = False):
    """pybind11 modules python/{tensors,sparse_matrices,periodic_homogenization} (src/python_bindings/bindings.cc)
    over the host classes; they link libmfem_b200.so like the CLIs do."""
    import sysconfig
    import pybind11
    os.makedirs(PY_DIR, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    src = os.path.join(REPO, "src", "python_bindings", "bindings.cc")
    h = hashlib.sha256()
    h.update(_host_stamp().encode())
    with open(src, "rb") as fh:
        h.update(fh.read())
    stamp = h.hexdigest()
    stamp_file = os.path.join(PY_DIR, "build.stamp")
    outs = [os.path.join(PY_DIR, name + suffix) for _, name in PY_MODULES]
    if not force and all(os.path.exists(o) for o in outs) and os.path.exists(stamp_file):
        if open(stamp_file).read().strip() == stamp:
            return outs
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include(), "-I" + os.path.join(REPO, "include")]
    procs = []
    for (define, name), out in zip(PY_MODULES, outs):
        cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-fvisibility=hidden", "-DBIND_" + define] + inc + \
              [src, os.path.join(REPO, "src", "host", "MeshIO.cc"), "-L" + LIBDIR, "-lmfem_b200",
               "-Wl,-rpath,$ORIGIN/../meshfem_b200/lib", "-o", out]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"pybind11 module {name} failed to build:\n{out}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return outs


def build_all(force: bool = False, verbose: bool = False):
    lib, host = build(force, verbose), build_host(force)
    build_python(force)
    return lib, host


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
