"""The synthetic-input generators of the bench (chemtensor_b200/workloads.py) against the reference's own generators."""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi, workloads


def _ref_mpo_matrix(ref, mpo):
    tensors = [mpo.site(i).to_dense() for i in range(mpo.nsites)]
    return workloads.mpo_to_matrix(tensors)


@pytest.mark.parametrize("model,L,params", [("xxz", 5, (1.0, 0.8, 0.1)), ("fermi_hubbard", 4, (1.0, 4.0, 0.3))])
def test_mpo_equals_reference_hamiltonian(ref, model, L, params):
    """Same operator as the reference's constructor (hamiltonian.c:102 / :240), compared as dense matrices."""
    tensors, qbonds, qsite = workloads.MODELS[model](L, *params)
    mine = workloads.mpo_to_matrix(tensors)
    theirs = _ref_mpo_matrix(ref, helpers.ref_mpo(ref, model, L, *params))
    assert np.array_equal(qsite, helpers.ref_mpo(ref, model, L, *params).qsite)
    assert np.max(np.abs(mine - theirs)) <= 1e-14
    assert np.max(np.abs(mine - mine.T)) == 0.0


def test_mpo_chain_obeys_quantum_numbers(ref):
    """Every entry of the generated MPO tensors sits in a conserving block (nothing is lost going block-sparse)."""
    for model, params in (("xxz", (1.0, 0.8, 0.1)), ("fermi_hubbard", (1.0, 4.0, 0.3))):
        tensors, qbonds, qsite = workloads.MODELS[model](6, *params)
        chain = workloads.mpo_chain(ref, model, 6, params)
        for i, t in enumerate(tensors):
            assert np.array_equal(chain.site(i).to_dense(), t)


@pytest.mark.parametrize("model,L,params,sector,D", [("xxz", 10, (1.0, 0.8, 0.1), 0, 20), ("fermi_hubbard", 8, (1.0, 4.0, 0.0), workloads.encode_qpair(8, 0), 50)])
def test_heff_operands_have_the_sweep_structure(eng, ref, model, L, params, sector, D):
    """Axis directions / dummy-leg quantum numbers equal those of real environments, and Heff on the synthetic operands
    agrees with the reference (1e-12 relative)."""
    mpo_r = helpers.ref_mpo(ref, model, L, *params)
    psi_r = helpers.ref_random_mps(ref, np.float64, L, mpo_r.qsite, sector, D, seed=1)
    rl = (cabi.BlockSparseTensor * L)()
    ref.compute_right_operator_blocks(psi_r.ptr, psi_r.ptr, mpo_r.ptr, rl)
    real_r = cabi.BST(ref, rl[L // 2])
    real_l = cabi.BST(ref)
    ref.create_dummy_operator_block_left(psi_r.site(0).ptr, psi_r.site(0).ptr, mpo_r.site(0).ptr, real_l.ptr)
    nxt = cabi.BST(ref)
    ref.contraction_operator_step_left(psi_r.site(0).ptr, psi_r.site(0).ptr, mpo_r.site(0).ptr, real_l.ptr, nxt.ptr)

    a, w, l, r = workloads.heff_operands(eng, model, L, params, sector, D, dtype=np.float64, seed=3)
    assert l.axis_dir == nxt.axis_dir and r.axis_dir == real_r.axis_dir
    assert np.array_equal(l.qnums[0], nxt.qnums[0]) and np.array_equal(r.qnums[3], real_r.qnums[3])
    ar, wr, lr, rr = (cabi.bst_clone(ref, x) for x in (a, w, l, r))
    be, br = cabi.BST(eng), cabi.BST(ref)
    eng.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, be.ptr)
    ref.apply_local_hamiltonian(ar.ptr, wr.ptr, lr.ptr, rr.ptr, br.ptr)
    helpers.assert_bst_close(be, br, 1e-12)
    for k in range(L):
        if k != L // 2:
            ref.delete_block_sparse_tensor(C.byref(rl[k]))


def test_flop_count_matches_engine_plans(eng):
    """bench.py's metadata-only flop count equals the engine's plan-time count (sum 2 m n k over the block GEMMs)."""
    from chemtensor_b200 import flops
    for model, L, params, sector, D, dtype in (("xxz", 12, (1.0, 0.8, 0.1), 0, 30, np.float64),
                                               ("fermi_hubbard", 8, (1.0, 4.0, 0.0), workloads.encode_qpair(8, 0), 40, np.complex128)):
        a, w, l, r = workloads.heff_operands(eng, model, L, params, sector, D, dtype=dtype, seed=5)
        ms, fl = C.c_double(0), C.c_double(0)
        assert eng.ctb_heff_benchmark(a.ptr, w.ptr, l.ptr, r.ptr, 0, 1, 0, C.byref(ms), C.byref(fl), None, None) == 0
        assert fl.value == flops.heff_flops(a, w, l, r)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_shard_partition_is_exact_and_balanced(world):
    """Plan-only view of the sharded matvec (no device work): the per-rank algorithmic flops of the three contractions add up to the
    single-rank count EXACTLY (the shards partition the block GEMMs column-wise: a checksum of checksums), every rank's result slice
    adds up to the full vector, and the partition (balanced in PADDED tile work, which is what a rank executes) keeps the algorithmic
    flops of the ranks within 25 % of the mean even at D=1024 on 8 ranks."""
    import bench
    eng = helpers.load("emu")
    a, w, l, r = bench.build_operands(eng, "fh_L32_D1024")
    info = (C.c_double * 8)()
    assert eng.ctb_heff_plan_info(a.ptr, w.ptr, l.ptr, r.ptr, 0, 1, info) == 0
    total_flops, total_entries = info[0], info[4]
    assert total_entries == a.num_elements()
    flops_p, entries_p = [], []
    for rank in range(world):
        assert eng.ctb_heff_plan_info(a.ptr, w.ptr, l.ptr, r.ptr, rank, world, info) == 0
        flops_p.append(info[0]); entries_p.append(info[4])
    assert sum(flops_p) == total_flops
    assert sum(entries_p) == total_entries
    assert max(flops_p) <= 1.25 * total_flops / world
