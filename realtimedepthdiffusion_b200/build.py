"""Build recipes for the native pieces (no build system beyond nvcc / gcc).

  librtdd.so            the product: sm_100a kernels + C ABI + reference-named C++ shims
  oracle/liboracle.so   CPU restatement (test infrastructure)
  oracle/_ref/libref.so the reference's own three .cu files, compiled UNMODIFIED from where
                        they lie under /root/reference (only when that tree is present; the
                        GPU box uses the prebuilt file that travels with the snapshot)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "realtimedepthdiffusion_b200", "csrc")
LIBDIR = os.path.join(ROOT, "realtimedepthdiffusion_b200", "lib")
LIBRTDD = os.path.join(LIBDIR, "librtdd.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIBORACLE = os.path.join(ORACLE_DIR, "liboracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
LIBREF = os.path.join(REF_DIR, "libref.so")
REFERENCE = "/root/reference"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def rtdd_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp"))]
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".h")]
    hdrs += [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include")))]
    return srcs, hdrs


def build_rtdd(force=False, verbose=False, extra=()):
    srcs, hdrs = rtdd_sources()
    if not force and not _stale(LIBRTDD, srcs + hdrs + [os.path.abspath(__file__)]):
        return LIBRTDD
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc()] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                              "-Xlinker", "-Bsymbolic", "-x", "cu",
                              "-I", os.path.join(ROOT, "include"), "-I", CSRC] + list(extra) + os.environ.get("RTDD_EXTRA_NVCC", "").split() + srcs + ["-o", LIBRTDD]
    _run(cmd, verbose)
    return LIBRTDD


def build_oracle(force=False, verbose=False):
    src = os.path.join(ORACLE_DIR, "depth_oracle.c")
    if not force and not _stale(LIBORACLE, [src]):
        return LIBORACLE
    cmd = ["gcc", "-O2", "-fopenmp", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", src, "-o", LIBORACLE, "-lm"]
    _run(cmd, verbose)
    return LIBORACLE


def build_ref(force=False, verbose=False):
    """Reference kernels, unmodified, default fp flags (what a user of the reference would get)."""
    srcs = [os.path.join(REFERENCE, "src", f) for f in ("GPUSolver.cu", "GPUImageProcessing.cu", "GPUDepthEffect.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return LIBREF if os.path.exists(LIBREF) else None
    if not force and not _stale(LIBREF, srcs):
        return LIBREF
    os.makedirs(REF_DIR, exist_ok=True)
    cmd = [_nvcc()] + ARCH + ["-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "-Xlinker", "-Bsymbolic", "-w",
                              "-I", os.path.join(REFERENCE, "include")] + srcs + ["-o", LIBREF]
    _run(cmd, verbose)
    return LIBREF


def build_all(force=False, verbose=False):
    out = {"librtdd": build_rtdd(force, verbose), "liboracle": build_oracle(force, verbose), "libref": build_ref(force, verbose)}
    return out


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
