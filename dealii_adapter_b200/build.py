"""Build recipes for the in-tree native libraries (sm_100a only; no JIT cache, no fallbacks).

  libgraftfem.so   CUDA kernels + C-ABI (include/graft_fem.h)           <- csrc/*.cu
  libgraft_host.so C++ host mirror of the reference classes + mesh       <- host/*.cc
The CPU oracle (oracle/liboracle.so) is test infrastructure and is built by oracle/Makefile.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
HOST = os.path.join(PKG_DIR, "host")
INCLUDE = os.path.join(REPO_DIR, "include")

LIB_CUDA = os.path.join(PKG_DIR, "libgraftfem.so")
LIB_HOST = os.path.join(PKG_DIR, "libgraft_host.so")
LIB_ORACLE = os.path.join(REPO_DIR, "oracle", "liboracle.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, log=None):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return res.stdout


def _sources(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_cuda(force=False):
    """One object per .cu (compiled in parallel, only when the source or a header changed), then
    one link: libgraftfem.so. Per-file ptxas -v output goes to csrc/build.log."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = _sources(CSRC, (".cu",))
    hdrs = _sources(CSRC, (".h", ".cuh")) + _sources(INCLUDE, (".h",))
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, os.path.basename(x)[:-3] + ".o") for x in srcs]
    todo = [(x, o) for x, o in zip(srcs, objs) if force or _newer(o, [x] + hdrs)]

    def compile_one(job):
        src, obj = job
        return _run([nvcc] + NVCC_FLAGS + ["-c", "-I", INCLUDE, "-I", CSRC, "-o", obj, src],
                    log=obj[:-2] + ".log")

    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(compile_one, todo))
    if todo or force or _newer(LIB_CUDA, objs):
        _run([nvcc, "-shared", "-o", LIB_CUDA] + objs + ["-lcudart", "-ldl"])
        with open(os.path.join(CSRC, "build.log"), "w") as f:
            for o in objs:
                lg = o[:-2] + ".log"
                if os.path.exists(lg):
                    f.write(open(lg).read())
    return LIB_CUDA


def build_host(force=False):
    """Mesh/DoF scaffolding only (no CUDA dependency): used by the Python binding, tests, oracle."""
    srcs = [os.path.join(HOST, "structured_mesh.cc"), os.path.join(HOST, "vtk_output.cc")]
    deps = srcs + [os.path.join(HOST, "structured_mesh.h"), os.path.join(HOST, "vtk_output.h"),
                   os.path.join(CSRC, "fe_basis.h")]
    if force or _newer(LIB_HOST, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", HOST, "-I", CSRC,
              "-o", LIB_HOST] + srcs)
    return LIB_HOST


def build_elasticity(force=False):
    """The drop-in host driver (C++ mirror of elasticity.cc + Solid/ElastoDynamics/Adapter) linked
    against libgraftfem.so; one executable per DIM like the reference's -DDIM build."""
    build_cuda()
    srcs = _sources(HOST, (".cc",))
    deps = srcs + _sources(HOST, (".h",)) + _sources(os.path.join(HOST, "adapter"), (".h",)) \
        + _sources(INCLUDE, (".h",)) + [LIB_CUDA, os.path.join(CSRC, "fe_basis.h")]
    outs = []
    for dim in (2, 3):
        exe = os.path.join(PKG_DIR, "elasticity_%dd" % dim)
        if force or _newer(exe, deps):
            _run(["g++", "-O2", "-std=c++17", "-Wall", "-DDIM=%d" % dim, "-I", INCLUDE, "-I", HOST,
                  "-I", CSRC, "-o", exe] + srcs + ["-L", PKG_DIR, "-lgraftfem", "-Wl,-rpath,$ORIGIN",
                                       "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
        outs.append(exe)
    return outs


def build_oracle(force=False):
    d = os.path.join(REPO_DIR, "oracle")
    if force:
        _run(["make", "-C", d, "clean"])
    _run(["make", "-C", d])
    return LIB_ORACLE


def build_oracle_ref():
    """oracle/_ref/: the pieces of the reference that compile without the real deal.II against
    oracle/ref_shim - the material, Postprocessor and Time headers in place, the cell-assembly
    structs of nonlinear_elasticity.cc and two statement blocks of linear_elasticity.cc cut out by
    marker at build time. Only in the build container; returns None where the reference tree does
    not exist (GPU box)."""
    if not os.path.isdir("/root/reference"):
        return None
    _run(["make", "-C", os.path.join(REPO_DIR, "oracle"), "ref"])
    return os.path.join(REPO_DIR, "oracle", "_ref", "ref_driver")


def build_all(force=False):
    return build_host(force), build_cuda(force), build_elasticity(force), build_oracle(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
