#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.  Run in the build container, where the reference tree is mounted read-only at
/root/reference (it does not exist on the GPU box; the .npz files are what travels).

1. The reference's own golden vectors for the hot path: the HDF5 fixtures its C test-suite loads
   are converted verbatim (datasets under 'ds/<name>', root attributes under 'at/<name>') with the minimal reader
   oracle/hdf5_v0.py -- data only, no reference source is copied.
2. Known answers produced by the UNMODIFIED compiled reference (oracle/_ref/libchemtensor_ref.so) on seeded inputs for
   the entry points that have no fixture of their own in the reference (apply_local_hamiltonian and the environment
   steps are only pinned indirectly there): operands and results of one two-site Heff application and
   of one left / right environment step on a small Fermi-Hubbard chain, plus per-sweep energies of short two-site
   DMRG runs on the BASELINE.json model families at small bond dimension.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF = os.environ.get("CTB_REFERENCE_TREE", "/root/reference")

FIXTURES = {
    "perf_dmrg_coeffs": "perf/perf_dmrg_coeffs.hdf5",      # integrals of BASELINE.json configs[0] (perf/perf_dmrg.c)
    "dmrg_twosite": "test/algorithm/data/test_dmrg_twosite.hdf5",
    "dmrg_singlesite": "test/algorithm/data/test_dmrg_singlesite.hdf5",
    "retained_bond_indices": "test/algorithm/data/test_retained_bond_indices.hdf5",
    "split_block_sparse_matrix_svd": "test/algorithm/data/test_split_block_sparse_matrix_svd.hdf5",
    "mpo_inner_product": "test/algorithm/data/test_mpo_inner_product.hdf5",
    "lanczos_iteration_d": "test/util/data/test_lanczos_iteration_d.hdf5",
    "lanczos_iteration_z": "test/util/data/test_lanczos_iteration_z.hdf5",
    "eigensystem_krylov_symmetric": "test/util/data/test_eigensystem_krylov_symmetric.hdf5",
    "eigensystem_krylov_hermitian": "test/util/data/test_eigensystem_krylov_hermitian.hdf5",
    "block_sparse_tensor_dot": "test/tensor/data/test_block_sparse_tensor_dot.hdf5",
    "block_sparse_tensor_qr": "test/tensor/data/test_block_sparse_tensor_qr.hdf5",
    "block_sparse_tensor_rq": "test/tensor/data/test_block_sparse_tensor_rq.hdf5",
    "block_sparse_tensor_svd": "test/tensor/data/test_block_sparse_tensor_svd.hdf5",
    "block_sparse_tensor_transpose": "test/tensor/data/test_block_sparse_tensor_transpose.hdf5",
    "block_sparse_tensor_reshape": "test/tensor/data/test_block_sparse_tensor_reshape.hdf5",
    "block_sparse_tensor_matricize_axis": "test/tensor/data/test_block_sparse_tensor_matricize_axis.hdf5",
    "block_sparse_tensor_slice": "test/tensor/data/test_block_sparse_tensor_slice.hdf5",
    "block_sparse_tensor_serialize": "test/tensor/data/test_block_sparse_tensor_serialize.hdf5",
    "block_sparse_tensor_multiply_pointwise_vector": "test/tensor/data/test_block_sparse_tensor_multiply_pointwise_vector.hdf5",
    "block_sparse_tensor_cyclic_partial_trace": "test/tensor/data/test_block_sparse_tensor_cyclic_partial_trace.hdf5",
    "mps_split_tensor_svd": "test/state/data/test_mps_split_tensor_svd.hdf5",
    "mps_orthonormalize_qr": "test/state/data/test_mps_orthonormalize_qr.hdf5",
}


def convert_fixtures():
    import hdf5_v0
    for name, rel in FIXTURES.items():
        ds, at = hdf5_v0.load(os.path.join(REF, rel))
        out = {f"ds/{k}": v for k, v in ds.items()}
        out.update({f"at/{k}": v for k, v in at.items()})
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **out)
        print(f"ref_{name}.npz: {len(ds)} datasets, {len(at)} attributes")


def bst_pack(prefix, t, out):
    """store a host block-sparse tensor: logical quantum numbers, directions and the serialised entries"""
    out[f"{prefix}/axis_dir"] = np.asarray(t.axis_dir, dtype=np.int32)
    for i, q in enumerate(t.qnums):
        out[f"{prefix}/qnums{i}"] = q
    out[f"{prefix}/entries"] = t.serialize()


def reference_known_answers():
    import helpers
    from chemtensor_b200 import cabi
    ref = helpers.load("ref")
    out = {}
    # --- one Heff application and one environment step each way, Fermi-Hubbard L=8, D<=48, float64 (complex128 is pinned by ref_dmrg_twosite.npz) ---
    for tag, dtype in (("d", np.float64),):
        L = 8
        mpo = helpers.ref_mpo(ref, "fermi_hubbard", L, 1.0, 4.0, 0.0)
        psi = helpers.ref_random_mps(ref, dtype, L, mpo.qsite, helpers.encode_qpair(L, 0), 48, seed=42)
        rl = (cabi.BlockSparseTensor * L)()
        ref.compute_right_operator_blocks(psi.ptr, psi.ptr, mpo.ptr, rl)
        i = L // 2 - 1
        l = cabi.BST(ref)
        ref.create_dummy_operator_block_left(psi.site(0).ptr, psi.site(0).ptr, mpo.site(0).ptr, l.ptr)
        for j in range(i):
            nl = cabi.BST(ref)
            ref.contraction_operator_step_left(psi.site(j).ptr, psi.site(j).ptr, mpo.site(j).ptr, l.ptr, nl.ptr)
            l = nl
        a, w, b = cabi.BST(ref), cabi.BST(ref), cabi.BST(ref)
        ref.mps_merge_tensor_pair(psi.site(i).ptr, psi.site(i + 1).ptr, a.ptr)
        ref.mpo_merge_tensor_pair(mpo.site(i).ptr, mpo.site(i + 1).ptr, w.ptr)
        r = cabi.BST(ref, rl[i + 1], owned=False)
        ref.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr)
        for nm, t in (("a", a), ("w", w), ("l", l), ("r", r), ("b", b)):
            bst_pack(f"heff_{tag}/{nm}", t, out)
        # environment steps from the same state
        ln = cabi.BST(ref)
        ref.contraction_operator_step_left(psi.site(i).ptr, psi.site(i).ptr, mpo.site(i).ptr, l.ptr, ln.ptr)
        rn = cabi.BST(ref)
        ref.contraction_operator_step_right(psi.site(i + 1).ptr, psi.site(i + 1).ptr, mpo.site(i + 1).ptr, r.ptr, rn.ptr)
        for nm, t in (("a_left", psi.site(i)), ("w_left", mpo.site(i)), ("l_next", ln), ("a_right", psi.site(i + 1)), ("w_right", mpo.site(i + 1)), ("r_next", rn)):
            bst_pack(f"env_{tag}/{nm}", t, out)
        for k in range(L):
            ref.delete_block_sparse_tensor(C.byref(rl[k]))
    # --- per-sweep energies of short two-site DMRG runs (the BASELINE.json model families, small bond dimension) ---
    for model, L, params, sector, D in (("xxz", 16, (1.0, 0.8, 0.1), 0, 32), ("fermi_hubbard", 8, (1.0, 4.0, 0.0), helpers.encode_qpair(8, 0), 48)):
        mpo = helpers.ref_mpo(ref, model, L, *params)
        psi = helpers.ref_random_mps(ref, np.float64, L, mpo.qsite, sector, D, seed=42)
        nsweeps = 3
        en = np.zeros(nsweeps); ent = np.zeros(L - 1)
        rc = ref.dmrg_twosite(mpo.ptr, nsweeps, 20, 1e-10, D, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        out[f"dmrg_{model}/params"] = np.array([L, D, nsweeps, 20, sector] + list(params), dtype=np.float64)
        out[f"dmrg_{model}/en_sweeps"] = en
        out[f"dmrg_{model}/entropy"] = ent
        out[f"dmrg_{model}/bond_dims"] = np.array(psi.bond_dims(), dtype=np.int64)
        print(f"dmrg_{model}: E = {en}")
    np.savez_compressed(os.path.join(HERE, "refrun_known_answers.npz"), **out)
    print(f"refrun_known_answers.npz: {len(out)} arrays")


if __name__ == "__main__":
    convert_fixtures()
    reference_known_answers()
