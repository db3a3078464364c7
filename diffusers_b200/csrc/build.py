"""Builds libb200diff.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_C")
LIB = os.path.join(OUT_DIR, "libb200diff.so")
SOURCES = ["lib.cu", "conv_gemm.cu", "attention.cu", "attention_pipe.cu", "attention64.cu", "norm.cu", "elementwise.cu", "peer.cu", "text_attention.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "b200_diffusion.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(HERE, s)
        obj = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, src):
            cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
            procs.append((s, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, cmd, p in procs:
        out, _ = p.communicate()
        log = os.path.join(OUT_DIR, s + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stdout.write(out)
    if failed:
        raise RuntimeError("nvcc failed (see above)")
    if procs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
