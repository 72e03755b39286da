"""TEST-ONLY: build the kernel sources on the SIMT emulator (simt_emu.h) with g++ so their logic can be
debugged in the GPU-less dev container.  The result is bound only by tests (explicit path)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "systemlevelcontrol.jl_b200", "csrc")
OUT = os.path.join(HERE, "libslsb200_emu.so")
SOURCES = ["sls_api.cu", "sls_gather.cu", "sls_solve.cu", "sls_solve_wide.cu", "sls_solve_huge.cu", "sls_solve_general.cu", "sls_scatter.cu", "sls_sim.cu", "sls_supports.cu"]


def build(force=False, opt="-O2"):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")] + [os.path.join(HERE, "simt_emu.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    objs = []
    for s in srcs:
        o = os.path.join(HERE, os.path.basename(s) + ".emu.o")
        if force or not os.path.exists(o) or any(os.path.getmtime(d) > os.path.getmtime(o) for d in [s] + deps[len(srcs):]):
            cmd = ["g++", opt, "-g", "-std=c++17", "-DSLS_EMU", "-fPIC", "-w", "-x", "c++", "-I", HERE, "-c", s, "-o", o]
            subprocess.run(cmd, check=True)
        objs.append(o)
    subprocess.run(["g++", "-shared", "-o", OUT, *objs], check=True)
    return OUT


if __name__ == "__main__":
    print(build())
