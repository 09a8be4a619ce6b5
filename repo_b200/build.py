"""Build librepo_b200.so (sm_100a) in-tree with nvcc.  `python -m repo_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "api.cu")
DEPS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))) + [
    os.path.join(os.path.dirname(HERE), "include", "repo_b200.h")]
OUT = os.path.join(HERE, "librepo_b200.so")
OUT_PROF = os.path.join(HERE, "librepo_b200_prof.so")   # -DRB_STAGE_CLOCK: in-kernel clock stamps (scripts/stage_clock.py)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False, profiling: bool = False) -> str:
    if not profiling and not force and up_to_date():
        return OUT
    out = OUT_PROF if profiling else OUT
    cmd = [NVCC] + FLAGS + (["-DRB_STAGE_CLOCK"] if profiling else []) + ["-o", out, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building librepo_b200.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, profiling="--profiling" in sys.argv))
