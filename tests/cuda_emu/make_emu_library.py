"""TEST INFRASTRUCTURE ONLY: builds tests/cuda_emu/_build/libgraftfem_emu.so, the library's own
sources (C-ABI, host glue, every kernel that needs neither TMA nor mbarriers) compiled with g++
against the CUDA stand-in of this directory, so that the C-ABI can be driven on a machine without a
GPU (tests/test_emulated_library.py). The .cu files are used as they are except for two textual
rewrites g++ cannot do without:
    kernel<<<grid, block, smem, stream>>>(args);  ->  gf_emu::launch4(grid, block, smem, stream, [&] { kernel(args); });
    extern __shared__ T name[];                   ->  T *name = reinterpret_cast<T *>(gf_emu::dynamic_smem());
Hardware-only parts are compiled out by the sources' own `#ifndef GF_CUDA_EMULATION` (tuned
neo-Hookean kernels -> generic kernels, TMA SpMV -> LDG SpMV) or replaced by emu_stubs.cpp
(single-launch coarsest-level solver -> multi-launch fallback; communicators:
GF_ERR_UNSUPPORTED). The product never loads this library; it has no CPU path."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "dealii_adapter_b200", "csrc")
OUT_DIR = os.environ.get("GF_EMU_BUILD_DIR", os.path.join(HERE, "_build"))
LIB = os.path.join(OUT_DIR, "libgraftfem_emu.so")
SOURCES = ["api.cu", "pattern.cu", "scatter.cu", "assemble_nl.cu", "assemble_lin.cu", "cg.cu",
           "reduce.cu", "vector_ops.cu", "constraints.cu", "direct.cu", "postprocess.cu",
           "fe_tables.cu", "spmv.cu", "operator.cu", "multigrid.cu", "matfree.cu", "comm.cu"]


def _matching(text, i, open_ch, close_ch):
    """index just after the bracket that closes text[i] == open_ch"""
    depth = 0
    for k in range(i, len(text)):
        if text[k] == open_ch:
            depth += 1
        elif text[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k + 1
    raise ValueError("unbalanced %s" % open_ch)


def rewrite(text):
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:]+)\s+(\w+)\s*\[\s*\]\s*;",
                  r"\1 *\2 = reinterpret_cast<\1 *>(gf_emu::dynamic_smem());", text)
    out, pos = [], 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            out.append(text[pos:])
            break
        # kernel expression: identifier with an optional template argument list, right before <<<
        j = i
        if text[j - 1] == ">":
            depth = 0
            while True:
                j -= 1
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        kernel = text[j:i]
        k = text.find(">>>", i)
        config = text[i + 3:k]
        a0 = k + 3
        while text[a0] in " \t\n\\":
            a0 += 1
        assert text[a0] == "(", text[i - 40:i + 80]
        a1 = _matching(text, a0, "(", ")")
        args = text[a0:a1]
        # the launch configuration always has four entries in this code base
        out.append(text[pos:j])
        out.append("gf_emu::launch4(%s, [&] { %s%s; })" % (config, kernel, args))
        pos = a1
    return "".join(out)


def build(force=False):
    os.makedirs(os.path.join(OUT_DIR, "gen"), exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
        [os.path.join(HERE, f) for f in ("cuda_runtime.h", "emu_stubs.cpp", "make_emu_library.py")] + \
        [os.path.join(HERE, "thrust", "thrust_emu.h"), os.path.join(ROOT, "include", "graft_fem.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    gen = []
    for f in SOURCES:
        dst = os.path.join(OUT_DIR, "gen", f[:-3] + "_emu.cpp")
        with open(os.path.join(CSRC, f)) as src, open(dst, "w") as o:
            o.write(rewrite(src.read()))
        gen.append(dst)
    gen.append(os.path.join(HERE, "emu_stubs.cpp"))
    flags = ["-O1", "-g", "-std=c++20", "-fPIC", "-pthread", "-DGF_CUDA_EMULATION",
             "-Wno-unknown-pragmas", "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include")]
    # GF_EMU_SANITIZE=address: "device" buffers are plain heap blocks, so AddressSanitizer sees
    # every out-of-bounds access of a kernel (run python with LD_PRELOAD=$(g++ -print-file-name=
    # libasan.so) ASAN_OPTIONS=detect_leaks=0)
    san = os.environ.get("GF_EMU_SANITIZE")
    link = []
    if san:
        flags += ["-fsanitize=" + san, "-fno-omit-frame-pointer"]
        link = ["-fsanitize=" + san]
    objs, procs = [], []
    for g in gen:
        obj = os.path.join(OUT_DIR, "gen", os.path.basename(g)[:-4] + ".o")
        objs.append(obj)
        procs.append((g, subprocess.Popen(["g++"] + flags + ["-c", g, "-o", obj],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for g, p in procs:
        log = p.communicate()[0]
        if p.returncode != 0:
            failed = True
            sys.stderr.write("---- %s\n%s\n" % (g, log[-6000:]))
    if failed:
        raise RuntimeError("emulation build failed")
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", LIB] + link + objs + ["-lrt", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
