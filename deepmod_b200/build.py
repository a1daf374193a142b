"""In-tree build of the CUDA library: nvcc -> deepmod_b200/libdeepmod_b200.so (sm_100a only)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdeepmod_b200.so")
SOURCES = ["dm_api.cu", "dm_features.cu", "dm_hist.cu", "dm_lstm_fp32.cu", "dm_lstm_tc.cu", "dm_cluster.cu", "dm_align.cu", "dm_signal.cu", "dm_reduce.cu", "dm_synth.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found: deepmod_b200 needs the CUDA toolkit to build its kernels")
    return exe


def stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "deepmod_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, jobs=None, out=None, extra=None):
    """Compile every .cu of the library for sm_100a and link the shared object.

    ``out`` / ``extra``: another output path and extra nvcc flags (tuning variants, see tools/tc_sweep.py)."""
    if out is None and not force and not stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build") if out is None else out + ".obj"
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src[:-3] + ".o")
        # DM_NVCC_EXTRA: extra flags for tuning sweeps (e.g. -DTC_NLUT=2); not used by the shipped build
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("DM_NVCC_EXTRA", "").split() + list(extra or []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        log, _ = p.communicate()
        if verbose and log:
            sys.stderr.write(log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, log))
        objs.append(obj)
    # NCCL is dlopen()ed by dm_reduce.cu (only its header is needed here), hence -ldl
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out or LIB] + objs + ["-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
