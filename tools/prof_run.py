"""Small drivers for ncu captures (one invocation of a path at a BASELINE config's size, nothing else):
    python tools/prof_run.py chain|grid|sim|supports [...]
Used as:  ncu --set full --clock-control none --import-source on -k regex:<kernel> -c <n> -o gpurun_out/<name> python tools/prof_run.py <mode>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slc_b200
from slc_b200 import workloads as W


def chain(syn, nx=100_000):
    A, B1, B2 = W.chain_plant(nx)
    Sx, Su = W.chain_supports(nx, 9, 29, 1.5)
    syn.set_plant(A, B2); syn.set_supports(Sx, Su)
    return nx


def main():
    modes = sys.argv[1:] or ["chain"]
    syn = slc_b200.Synthesizer()
    for mode in modes:
        if mode == "chain":
            chain(syn); syn.plan(None); syn.solve_device()
            print("chain", syn.timing())
        elif mode == "grid":
            side = int(os.environ.get("GRID_SIDE", "300"))
            A, B1, B2 = W.grid_plant(side, side)
            Sx, Su = W.grid_supports(side, side, 4, 20, 1.5)
            syn.set_plant(A, B2); syn.set_supports(Sx, Su); syn.plan(None); syn.solve_device()
            print("grid", syn.timing())
        elif mode == "sim":
            nx = chain(syn); syn.plan(None); syn.solve_device()
            rng = np.random.default_rng(0)
            js = rng.integers(0, nx, 4096); ts = rng.integers(1, 250 - 29, 4096)
            syn.simulate(js, ts, 250)
            print("sim", syn.timing())
        elif mode == "supports":
            A, B1, B2 = W.chain_plant(100_000)
            syn.set_plant(A, B2)
            for _ in range(2):
                syn.make_supports(9, 29, 1.5)
            print("supports", syn.timing())
    syn.close()


if __name__ == "__main__":
    main()
