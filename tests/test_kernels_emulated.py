"""Kernel LOGIC tests without a GPU: the very same csrc/*.cu sources compiled by g++ on the test-only SIMT
emulator (tests/emu).  This is a debugging aid like compute-sanitizer — it proves indexing / algorithm
logic, not the CUDA build; the parity tests proper are tests/test_parity_gpu.py (`-m gpu`)."""
import numpy as np

import parity_cases as pc
from oracle import sls_oracle as O


def test_emu_reduction_fixture(emu_api):
    pc.check_reduction_fixture(emu_api)


def test_emu_reduction_output_feedback_branch(emu_api):
    pc.check_reduction_of(emu_api)


def test_emu_readme_chain_subset_vs_golden(emu_api):
    worst, info = pc.check_golden(emu_api, "readme_chain_59", O.chain_plant(59), cols=[0, 1, 5, 29, 30, 57, 58])
    assert worst < 1e-11


def test_emu_local_blocks_bit_exact(emu_api):
    A, B1, B2, S = pc.chain(59)
    pc.check_local_blocks(emu_api, A, B1, B2, S, [0, 3, 29, 58])


def test_emu_grid_9x9_subset(emu_api):
    worst, _ = pc.check_golden(emu_api, "grid_9x9_full_d2_T8", O.grid_plant(9, 9, actuation="full"), cols=[0, 4, 40, 44, 80])
    assert worst < 1e-11


def test_emu_wide_solver_grid_13x13_subset(emu_api):
    """n = 36 and n = 61 columns: the one-CTA-per-column kernel of the large size classes (sls_solve_wide.cu)."""
    from conftest import PHI_FLOOR
    worst, info = pc.check_golden(emu_api, "grid_13x13_full_d4_T20_subset", O.grid_plant(13, 13, actuation="full"), cols=[6, 84],
                                  floor=PHI_FLOOR)
    assert info["n"].tolist() == [36, 61] and worst < 1e-10


def test_emu_error_paths(emu_api):
    pc.check_error_paths(emu_api)


def test_emu_gather_fast_and_general_paths(emu_api):
    assert pc.check_gather_paths(emu_api) < 1e-9


def test_emu_large_local_problems_n_above_64(emu_api):
    assert pc.check_large_local_problems(emu_api) < 1e-10


def test_emu_masks_and_flags(emu_api):
    pc.check_masks_and_flags(emu_api)
    pc.check_unsorted_plan(emu_api)


def test_emu_closed_loop_simulation(emu_api):
    worst, tm = pc.check_simulation(emu_api, 100, [(50, 5), (0, 1), (99, 60), (7, 119)], steps=120)
    assert tm["sim_window"] < 100        # genuinely windowed: 2-level closure (radius 18) + one hop = 39 states


def test_emu_device_supports(emu_api):
    pc.check_device_supports(emu_api, O.chain_plant(40), 9, 29, 1.5)
    pc.check_device_supports(emu_api, O.grid_plant(6, 7, actuation="full"), 2, 6, 1.5)


def test_emu_pipelined_one_shot_equals_three_call(emu_api):
    A, B1, B2, S = pc.chain(24, d=3, T=8)
    pc.check_pipelined(emu_api, (A, B1, B2), S, chunks=[1, 5])


def test_emu_pipelined_rowval_from_host_or_device(emu_api):
    pc.check_pipelined_rowval_sources(emu_api)


def test_emu_window_shard_form(emu_api):
    A, B1, B2, S = pc.chain(30, d=3, T=8)
    pc.check_window(emu_api, (A, B1, B2), S, ranges=[(0, 11), (9, 22), (20, 30)], sub_reach=5)


def test_emu_weighted_cost_and_groups(emu_api):
    pc.check_weighted_cost(emu_api, cols=(3, 39, 14))


def test_emu_general_cost_and_coupled_groups(emu_api):
    """kernel 2g (sls_solve_general.cu) vs the oracle's null-space restatement."""
    assert pc.check_coupled_groups_identity_cost(emu_api)[0] < 1e-10
    assert pc.check_general_cost(emu_api) < 1e-10
    pc.check_general_error_paths(emu_api)


def test_emu_pipelined_ring_plant_long_range_coupling(emu_api):
    """A ring (A[0,Nx-1], A[Nx-1,0] != 0): the first column chunk reaches the last support columns and vice versa, so the
    copy-in / plan dependency logic of sls_synthesize degenerates to waiting for every upload; results must not change."""
    import scipy.sparse as sp
    Nx = 24
    A, B1, B2 = O.chain_plant(Nx)
    A = sp.lil_matrix(A); A[0, Nx - 1] = -0.2; A[Nx - 1, 0] = 0.2; A = sp.csc_matrix(A)
    S = O.make_supports(A, B2, 3, 8, 1.5)
    pc.check_pipelined(emu_api, (A, B1, B2), S, chunks=[4])


def test_emu_random_banded_plants_vs_oracle(emu_api):
    worst, nclean, nflag = pc.check_random_banded(emu_api, seeds=(0, 5), cols=(0, 5, 24, 47))
    assert nclean >= 6 and nflag >= 1 and worst < 1e-9
