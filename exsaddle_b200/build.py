"""In-tree build of libexsaddle_b200.so with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libexsaddle_b200.so")
SOURCES = ["xsb_api.cu", "xsb_fe.cu", "xsb_spmv.cu", "xsb_vec.cu", "xsb_mg.cu", "xsb_ilu.cu", "xsb_ksp.cu", "xsb_mf.cu", "xsb_mf1p.cu", "xsb_mfull.cu", "xsb_mmg.cu", "xsb_fs.cu", "xsb_asm.cu", "xsb_grad.cu", "xsb_io.cu", "xsb_comm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin", "/usr/bin/g++",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(out, deps):
    return (not os.path.exists(out)) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(verbose=False, force=False, ptxas_info=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, "xsb.h"), os.path.join(HERE, "..", "include", "exsaddle_b200.h"), os.path.abspath(__file__)]
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s); obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, rc, out in ex.map(run, jobs):
            if verbose or rc:
                sys.stderr.write(" ".join(cmd) + "\n" + out)
            if rc:
                raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, ptxas_info="--ptxas" in sys.argv))
