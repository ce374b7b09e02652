"""In-tree build of libgpfo.so (hand-written CUDA for sm_100a) with plain nvcc.

    python -m gpflowopt_b200.build [--force] [--verbose]

The shared library lands next to this file (gpflowopt_b200/libgpfo.so); it is git-ignored but travels to the
GPU box with the repo snapshot.  There is no CPU fallback: if the library is missing the package fails loudly.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libgpfo.so")
SOURCES = ["api.cu", "factor.cu", "contract.cu", "contract2.cu", "contract2_aux.cu", "thin.cu", "contract5.cu", "contract5_grad.cu", "hvpoi.cu", "grad.cu", "pareto_host.cu"]
# Kernels that were measured and rejected (2-CTA cluster, warp-specialised, bulk-copy re-reads) and the profiling builds
# (ablations, phase counters) are compiled only with GPFO_BUILD_EXPERIMENTS=1; the production library holds reachable code only.
EXPERIMENT_SOURCES = ["contract2_prof.cu", "contract3.cu", "contract4.cu"]
EXPERIMENTS = os.environ.get("GPFO_BUILD_EXPERIMENTS", "0") not in ("", "0")
HEADERS = [os.path.join(CSRC, "gpfo_internal.cuh"), os.path.join(CSRC, "contract_common.cuh"), os.path.join(CSRC, "contract2_kernel.cuh"),
           os.path.join(CSRC, "contract5_kernel.cuh"), os.path.join(CSRC, "contract5_mma.inc"), os.path.join(CSRC, "contract5_bwd_kernel.cuh"),
           os.path.join(CSRC, "ndtr_fast.cuh"), os.path.join(HERE, "..", "include", "gpfo.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--expt-extended-lambda"] + (["-DGPFO_BUILD_EXPERIMENTS=1"] if EXPERIMENTS else [])
STAMP = os.path.join(OBJ, ".flavour")


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libgpfo.so cannot be built")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(args):
    src, obj, verbose = args
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    flavour = "experiments" if EXPERIMENTS else "production"
    if not os.path.exists(STAMP) or open(STAMP).read() != flavour:     # objects of the other flavour are stale
        force = True
        with open(STAMP, "w") as f:
            f.write(flavour)
    sources = [s for s in SOURCES + (EXPERIMENT_SOURCES if EXPERIMENTS else []) if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + HEADERS):
            jobs.append((src, obj, verbose))
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for log in ex.map(_compile, jobs):
            if verbose and log:
                print(log)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
