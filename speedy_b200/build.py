"""Build libspeedy_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m speedy_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# developer builds (A/B variants, -DK4_TIMING) can go elsewhere; see SPEEDY_B200_LIB in __init__.py
OUT = os.environ.get("SPEEDY_B200_BUILD_OUT") or os.path.join(HERE, "libspeedy_b200.so")
BUILD = os.environ.get("SPEEDY_B200_BUILD_DIR") or os.path.join(HERE, "_build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O2,-Wall", "-I" + CSRC]
COMMON += os.environ.get("SPEEDY_B200_EXTRA_FLAGS", "").split()  # developer A/B builds, e.g. -DK1_RUN=7
if os.environ.get("SPEEDY_K4_TIMING"):
    COMMON.append("-DK4_TIMING")  # developer build: per-phase cycle counters in k4_sonic

# (source, extra flags).  The recurrence and Sonic kernels must round exactly as
# the reference's C does, so they are built without FMA contraction.
UNITS = [
    ("k1_spectral.cu", []),
    ("k1_dft16.cu", []),
    ("k2_tension.cu", ["--fmad=false"]),
    ("k4_sonic.cu", ["--fmad=false"]),
    ("k4_splice.cu", ["--fmad=false"]),
    ("k4_chain16.cu", ["--fmad=false"]),
    ("batch.cu", ["--fmad=false"]),
    ("sonic_api.cpp", []),
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "speedy_b200.h"))
    objs = []
    for src, extra in UNITS:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
    if force or _stale(OUT, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    # the speedy_wave-compatible command-line tool (host C++ only, links the library)
    tool_src = os.path.join(HERE, "..", "tools", "speedy_wave.cpp")
    tool = os.path.join(HERE, "speedy_wave")
    if os.path.exists(tool_src) and not os.environ.get("SPEEDY_B200_BUILD_OUT") and (force or _stale(tool, [tool_src, OUT] + headers)):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(HERE, "..", "include"), tool_src,
               "-L" + HERE, "-lspeedy_b200", "-Wl,-rpath,$ORIGIN", "-o", tool]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    # config 5 through the drop-in API: many pooled sonicStream handles fed 10 ms chunks (host C++ only)
    sb_src = os.path.join(HERE, "..", "tools", "stream_bench.cpp")
    sb_tool = os.path.join(HERE, "stream_bench")
    if os.path.exists(sb_src) and not os.environ.get("SPEEDY_B200_BUILD_OUT") and (force or _stale(sb_tool, [sb_src, OUT] + headers)):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", sb_src, "-L" + HERE, "-lspeedy_b200",
               "-Wl,-rpath,$ORIGIN", "-o", sb_tool]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    # the evaluation tools (DTW / Teager / slopes), host C++ only
    eval_src = os.path.join(HERE, "..", "tools", "eval_tools.cpp")
    eval_lib = os.path.join(HERE, "libspeedy_eval.so")
    eval_hdr = os.path.join(HERE, "..", "include", "speedy_eval.h")
    if os.path.exists(eval_src) and (force or _stale(eval_lib, [eval_src, eval_hdr])):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-ffp-contract=off",
               "-I" + os.path.join(HERE, "..", "include"), eval_src, "-o", eval_lib]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
