"""SU(2)-symmetric path (SURVEY 8(a) last row, 8(f) rank 2, BASELINE configs[4]) against the unmodified reference (oracle/_ref).

Mirrors the reference's own fixture-free SU(2) tests (test/algorithm/test_su2_chain_ops.c, test_su2_dmrg.c, test/tensor/test_su2_tensor.c):
inputs from the reference's generators, the same host structs through the engine's C entry points.  Bars: trees, irreducible lists,
degeneracy dimensions and charge-sector tables bit-exact; degeneracy-tensor entries 1e-12 relative; sweep energies 1e-10
(the tests assert tighter), entropies 1e-10.  'emu' = host logic on the CPU test double (runs without a GPU), 'cuda' = the product.
"""
import ctypes as C

import numpy as np
import pytest

import helpers
import su2_helpers as S

KINDS = ["emu", pytest.param("cuda", marks=pytest.mark.gpu)]
T = S.SU2Tensor


def _envs(r, psi, mpo, L):
    rl = (T * L)()
    r.su2_compute_right_operator_blocks(C.byref(psi), C.byref(psi), C.byref(mpo), rl)
    lb = T()
    r.su2_create_dummy_operator_block_left(1, C.byref(lb))
    lbs = [lb]
    for i in range(L - 1):
        nr = T()
        r.su2_contraction_operator_step_left(C.byref(psi.a[i]), C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[-1]), C.byref(nr))
        lbs.append(nr)
    return lbs, rl


def test_recoupling_coefficients_equal_reference_tables():
    """Racah-formula evaluation against the reference's generated tables over their whole range (ja, jb, jc <= 5), incl. the zero pattern"""
    r, e = S.ref(), S.engine("emu")
    worst = 0.0
    for ja in range(6):
        for jb in range(6):
            for jc in range(6):
                for js in range((ja + jb + jc) % 2, ja + jb + jc + 1, 2):
                    for je in range(abs(ja - jb), ja + jb + 1, 2):
                        for jf in range(abs(jb - jc), jb + jc + 1, 2):
                            a = r.su2_recoupling_coefficient(ja, jb, jc, js, je, jf)
                            b = e.su2_recoupling_coefficient(ja, jb, jc, js, je, jf)
                            assert (a == 0) == (b == 0), (ja, jb, jc, js, je, jf, a, b)
                            worst = max(worst, abs(a - b))
    assert worst < 5e-16


@pytest.mark.parametrize("kind", KINDS)
def test_contract_simple_and_fmove(kind):
    r, e = S.ref(), S.engine(kind)
    psi = S.random_mps(6, [1], [0, 1], 0, 5, 11, 41, scale=2.0)
    mpo = S.heisenberg_mpo(6, 0.7)
    for i in range(5):
        x, y = T(), T()
        ia, ib = (C.c_int * 1)(2), (C.c_int * 1)(0)
        r.su2_tensor_contract_simple(C.byref(psi.a[i]), ia, C.byref(psi.a[i + 1]), ib, 1, C.byref(x))
        e.su2_tensor_contract_simple(C.byref(psi.a[i]), ia, C.byref(psi.a[i + 1]), ib, 1, C.byref(y))
        S.assert_same_su2(y, x, 1e-13)
        ax = x.tree.tree_split.contents.c[0].contents.i_ax
        fx, fy = T(), T()
        r.su2_tensor_fmove(C.byref(x), ax, C.byref(fx))
        e.su2_tensor_fmove(C.byref(x), ax, C.byref(fy))
        S.assert_same_su2(fy, fx, 1e-13)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("model", ["heisenberg", "fermi_hubbard"])
def test_environment_steps_and_local_hamiltonian(kind, model):
    r, e = S.ref(), S.engine(kind)
    L = 6
    if model == "heisenberg":
        mpo = S.heisenberg_mpo(L, 1.3)
        psi = S.random_mps(L, [1], [0, 1], 0, 5, 9, 83, scale=3.0)
    else:
        mpo = S.fermi_hubbard_mpo(L, 1.0, 4.0, 0.3)
        # site: j = 0 twice (empty, doubly occupied), j = 1/2 once
        psi = S.random_mps(L, [0, 1], [2, 1], 0, 3, 7, 87, scale=2.0)
    rl_ref, rl_eng = (T * L)(), (T * L)()
    r.su2_compute_right_operator_blocks(C.byref(psi), C.byref(psi), C.byref(mpo), rl_ref)
    e.su2_compute_right_operator_blocks(C.byref(psi), C.byref(psi), C.byref(mpo), rl_eng)
    for i in range(L):
        S.assert_same_su2(rl_eng[i], rl_ref[i], 1e-12)
    lb_ref, lb_eng = T(), T()
    r.su2_create_dummy_operator_block_left(1, C.byref(lb_ref))
    e.su2_create_dummy_operator_block_left(1, C.byref(lb_eng))
    S.assert_same_su2(lb_eng, lb_ref, 0.0)
    rb_ref, rb_eng = T(), T()
    r.su2_create_dummy_operator_block_right(1, 3, C.byref(rb_ref))
    e.su2_create_dummy_operator_block_right(1, 3, C.byref(rb_eng))
    S.assert_same_su2(rb_eng, rb_ref, 0.0)
    lbs = [lb_ref]
    for i in range(L - 1):
        nr, ne = T(), T()
        r.su2_contraction_operator_step_left(C.byref(psi.a[i]), C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[-1]), C.byref(nr))
        e.su2_contraction_operator_step_left(C.byref(psi.a[i]), C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[-1]), C.byref(ne))
        S.assert_same_su2(ne, nr, 1e-12)
        lbs.append(nr)
    for i in range(L):
        br, be = T(), T()
        r.su2_apply_local_hamiltonian(C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[i]), C.byref(rl_ref[i]), C.byref(br))
        e.su2_apply_local_hamiltonian(C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[i]), C.byref(rl_ref[i]), C.byref(be))
        S.assert_same_su2(be, br, 1e-12)
    a, b = C.c_double(), C.c_double()
    r.su2_mpo_inner_product(C.byref(psi), C.byref(mpo), C.byref(psi), C.byref(a))
    e.su2_mpo_inner_product(C.byref(psi), C.byref(mpo), C.byref(psi), C.byref(b))
    assert abs(a.value - b.value) <= 1e-12 * max(1.0, abs(a.value))


@pytest.mark.parametrize("kind", KINDS)
def test_pair_form_equals_merged_reference(kind):
    """two site operators one after the other on the 4-leg tensor == the reference's merged pair tensor on the fused 3-leg tensor"""
    r, e = S.ref(), S.engine(kind)
    e.ctb_su2_apply_local_hamiltonian_pair.restype = None
    e.ctb_su2_apply_local_hamiltonian_pair.argtypes = [C.POINTER(T)] * 6
    L = 6
    mpo = S.heisenberg_mpo(L, 1.1)
    psi = S.random_mps(L, [1], [0, 1], 0, 5, 9, 85, scale=3.0)
    lbs, rl = _envs(r, psi, mpo, L)
    for i in range(L - 1):
        a2, am, h2, bm, b2, bf = T(), T(), T(), T(), T(), T()
        r.su2_mps_contract_tensor_pair(C.byref(psi.a[i]), C.byref(psi.a[i + 1]), C.byref(a2))
        r.su2_mps_merge_tensor_pair(C.byref(psi.a[i]), C.byref(psi.a[i + 1]), C.byref(am))
        r.su2_mpo_merge_tensor_pair(C.byref(mpo.a[i]), C.byref(mpo.a[i + 1]), C.byref(h2))
        r.su2_apply_local_hamiltonian(C.byref(am), C.byref(h2), C.byref(lbs[i]), C.byref(rl[i + 1]), C.byref(bm))
        e.ctb_su2_apply_local_hamiltonian_pair(C.byref(a2), C.byref(mpo.a[i]), C.byref(mpo.a[i + 1]), C.byref(lbs[i]), C.byref(rl[i + 1]), C.byref(b2))
        r.su2_tensor_fuse_axes(C.byref(b2), 1, 2, C.byref(bf))
        S.assert_same_su2(bf, bm, 1e-12)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("mode", [0, 1])
def test_mps_orthonormalize(kind, mode):
    r, e = S.ref(), S.engine(kind)
    L = 7
    mpo = S.heisenberg_mpo(L, 1.3)
    psi = S.random_mps(L, [1], [0, 1], 1, 5, 27, 81, scale=14.0)
    p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
    n1 = r.su2_mps_orthonormalize_qr(C.byref(p1), mode)
    n2 = e.su2_mps_orthonormalize_qr(C.byref(p2), mode)
    assert abs(n1 - n2) <= 1e-12 * n1
    assert r.su2_mps_is_consistent(C.byref(p2))
    a, b = C.c_double(), C.c_double()
    r.su2_mpo_inner_product(C.byref(p1), C.byref(mpo), C.byref(p1), C.byref(a))
    r.su2_mpo_inner_product(C.byref(p2), C.byref(mpo), C.byref(p2), C.byref(b))
    assert abs(a.value - b.value) <= 1e-12 * abs(a.value)
    # same bond structure
    for i in range(L):
        for ax in (0, 2):
            ja = [p1.a[i].outer_irreps[ax].jlist[k] for k in range(p1.a[i].outer_irreps[ax].num)]
            jb = [p2.a[i].outer_irreps[ax].jlist[k] for k in range(p2.a[i].outer_irreps[ax].num)]
            assert ja == jb
            assert [p1.a[i].dim_degen[ax][j] for j in ja] == [p2.a[i].dim_degen[ax][j] for j in jb]
    # the state itself: same vector (the gauge of the isometries drops out)
    v1, v2 = T(), T()
    r.su2_mps_to_statevector(C.byref(p1), C.byref(v1))
    r.su2_mps_to_statevector(C.byref(p2), C.byref(v2))
    assert r.su2_tensor_allclose(C.byref(v1), C.byref(v2), 1e-11)


@pytest.mark.parametrize("kind", KINDS)
def test_su2_dmrg_singlesite(kind):
    """the reference's test_su2_dmrg_singlesite (test/algorithm/test_su2_dmrg.c:15-125): same inputs, energies against the reference's run"""
    r, e = S.ref(), S.engine(kind)
    L, ns = 7, 3
    mpo = S.heisenberg_mpo(L, 1.3)
    psi = S.random_mps(L, [1], [0, 1], 1, 5, 27, 81, scale=14.0)
    p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
    e1, e2 = (C.c_double * ns)(), (C.c_double * ns)()
    assert r.su2_dmrg_singlesite(C.byref(mpo), ns, 5, C.byref(p1), e1) == 0
    assert e.su2_dmrg_singlesite(C.byref(mpo), ns, 5, C.byref(p2), e2) == 0
    assert np.allclose(list(e1), list(e2), rtol=0, atol=1e-11)
    assert r.su2_mps_is_consistent(C.byref(p2))
    b = C.c_double()
    r.su2_mpo_inner_product(C.byref(p2), C.byref(mpo), C.byref(p2), C.byref(b))
    assert abs(b.value - e2[ns - 1]) <= 1e-12 * abs(b.value)
    # exact ground state of the 7-site chain (dense diagonalisation of the spin-1/2 Heisenberg Hamiltonian)
    assert abs(e2[ns - 1] - _heisenberg_ground_state(L, 1.3)) < 1e-11


def _heisenberg_ground_state(L: int, J: float) -> float:
    sx = np.array([[0, 0.5], [0.5, 0]]); sy = np.array([[0, -0.5j], [0.5j, 0]]); sz = np.diag([0.5, -0.5])
    H = np.zeros((2 ** L, 2 ** L), dtype=complex)
    for i in range(L - 1):
        for s in (sx, sy, sz):
            H += J * np.kron(np.kron(np.eye(2 ** i), np.kron(s, s)), np.eye(2 ** (L - i - 2)))
    return float(np.linalg.eigvalsh(H)[0])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("tol,max_vdim", [(1e-5, 1000), (0.0, 12), (1e-3, 16)])
def test_su2_dmrg_twosite(kind, tol, max_vdim):
    """the reference's test_su2_dmrg_twosite (test/algorithm/test_su2_dmrg.c:128-): energies, entropies and bond structure against the reference"""
    r, e = S.ref(), S.engine(kind)
    L, ns = 7, 2
    mpo = S.heisenberg_mpo(L, 1.1)
    psi = S.random_mps(L, [1], [0, 1], 1, 5, 27, 82, scale=14.0)
    p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
    e1, e2 = (C.c_double * ns)(), (C.c_double * ns)()
    s1, s2 = (C.c_double * (L - 1))(), (C.c_double * (L - 1))()
    assert r.su2_dmrg_twosite(C.byref(mpo), ns, 5, tol, max_vdim, C.byref(p1), e1, s1) == 0
    assert e.su2_dmrg_twosite(C.byref(mpo), ns, 5, tol, max_vdim, C.byref(p2), e2, s2) == 0
    assert np.allclose(list(e1), list(e2), rtol=0, atol=1e-10), (list(e1), list(e2))
    assert np.allclose(list(s1), list(s2), rtol=0, atol=1e-9), (list(s1), list(s2))
    assert r.su2_mps_is_consistent(C.byref(p2))
    for i in range(L):
        ja = [p1.a[i].outer_irreps[2].jlist[k] for k in range(p1.a[i].outer_irreps[2].num)]
        jb = [p2.a[i].outer_irreps[2].jlist[k] for k in range(p2.a[i].outer_irreps[2].num)]
        assert ja == jb
        assert [p1.a[i].dim_degen[2][j] for j in ja] == [p2.a[i].dim_degen[2][j] for j in jb]
    b = C.c_double()
    r.su2_mpo_inner_product(C.byref(p2), C.byref(mpo), C.byref(p2), C.byref(b))
    assert abs(b.value - e2[ns - 1]) <= 1e-11 * abs(b.value)


@pytest.mark.parametrize("kind", KINDS)
def test_su2_dmrg_twosite_fermi_hubbard_longer_chain(kind):
    """a chain whose bonds grow through the split (L = 12, two physical irreducible sectors), against the reference"""
    r, e = S.ref(), S.engine(kind)
    L, ns = 10, 2
    mpo = S.fermi_hubbard_mpo(L, 1.0, 4.0, 1.5)
    psi = S.random_mps(L, [0, 1], [2, 1], 0, 3, 6, 91, scale=3.0)
    p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
    e1, e2 = (C.c_double * ns)(), (C.c_double * ns)()
    s1, s2 = (C.c_double * (L - 1))(), (C.c_double * (L - 1))()
    assert r.su2_dmrg_twosite(C.byref(mpo), ns, 8, 1e-8, 60, C.byref(p1), e1, s1) == 0
    assert e.su2_dmrg_twosite(C.byref(mpo), ns, 8, 1e-8, 60, C.byref(p2), e2, s2) == 0
    assert np.allclose(list(e1), list(e2), rtol=0, atol=1e-10), (list(e1), list(e2))
    assert np.allclose(list(s1), list(s2), rtol=0, atol=1e-8)
    assert r.su2_mps_is_consistent(C.byref(p2))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("copy_tree_left", [True, False])
def test_su2_tensor_svd_and_renormalized_entries(kind, copy_tree_left):
    """su2_tensor_svd (reference test/tensor/test_su2_tensor.c, test_su2_tensor_svd): structure bit-exact, singular values, isometries, reconstruction"""
    r, e = S.ref(), S.engine(kind)
    r.su2_tensor_fuse_axes_add_auxiliary.restype = None
    r.su2_tensor_fuse_axes_add_auxiliary.argtypes = [C.POINTER(T), C.c_int, C.c_int, C.POINTER(T)]
    r.su2_tensor_svd.restype = C.c_int
    r.su2_tensor_is_isometry.restype = C.c_bool
    r.su2_tensor_is_isometry.argtypes = [C.POINTER(T), C.c_double, C.c_bool]
    from chemtensor_b200 import cabi
    sig = [C.POINTER(T), C.c_bool, C.POINTER(T), C.POINTER(cabi.DenseTensor), C.POINTER(C.POINTER(C.c_int)), C.POINTER(T)]
    r.su2_tensor_svd.argtypes = sig
    e.su2_tensor_svd.restype = C.c_int
    e.su2_tensor_svd.argtypes = sig
    psi = S.random_mps(6, [1], [0, 1], 0, 5, 13, 57, scale=2.0)
    for i in (1, 2, 3):
        a = T()
        r.su2_tensor_fuse_axes_add_auxiliary(C.byref(psi.a[i]), 0, 1, C.byref(a))
        ur, vr, ue, ve = T(), T(), T(), T()
        sr, se = cabi.DenseTensor(), cabi.DenseTensor()
        mr, me = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
        assert r.su2_tensor_svd(C.byref(a), copy_tree_left, C.byref(ur), C.byref(sr), C.byref(mr), C.byref(vr)) == 0
        assert e.su2_tensor_svd(C.byref(a), copy_tree_left, C.byref(ue), C.byref(se), C.byref(me), C.byref(ve)) == 0
        n = sr.dim[0]
        assert se.dim[0] == n
        s1 = np.ctypeslib.as_array(C.cast(sr.data, C.POINTER(C.c_double)), shape=(n,)).copy()
        s2 = np.ctypeslib.as_array(C.cast(se.data, C.POINTER(C.c_double)), shape=(n,)).copy()
        assert np.allclose(s1, s2, rtol=0, atol=1e-13 * s1.max())
        assert [mr[k] for k in range(n)] == [me[k] for k in range(n)]
        S.assert_same_su2(ue, ur, 1e300)      # structure only: the factors are unique up to signs
        S.assert_same_su2(ve, vr, 1e300)
        assert r.su2_tensor_is_isometry(C.byref(ue), 1e-12, False)
        assert r.su2_tensor_is_isometry(C.byref(ve), 1e-12, True)
        off = 0
        for c in range(S.sectors(a).shape[0]):
            U, V, A = S.degensor(ue, c), S.degensor(ve, c), S.degensor(a, c)
            k = U.shape[1]
            assert np.allclose((U * s2[off:off + k]) @ V, A, rtol=0, atol=1e-12 * max(1.0, np.abs(A).max()))
            off += k
    # renormalised (de)serialisation: same packed vector as the reference, round trip
    e.su2_tensor_num_elements_degensors.restype = C.c_int64
    e.su2_tensor_num_elements_degensors.argtypes = [C.POINTER(T)]
    for lib in (r, e):
        lib.su2_tensor_serialize_renormalized_entries.restype = None
        lib.su2_tensor_serialize_renormalized_entries.argtypes = [C.POINTER(T), C.c_void_p]
        lib.su2_tensor_deserialize_renormalized_entries.restype = None
        lib.su2_tensor_deserialize_renormalized_entries.argtypes = [C.POINTER(T), C.c_void_p]
    t = psi.a[2]
    n = e.su2_tensor_num_elements_degensors(C.byref(t))
    v1, v2 = np.zeros(n), np.zeros(n)
    r.su2_tensor_serialize_renormalized_entries(C.byref(t), v1.ctypes.data)
    e.su2_tensor_serialize_renormalized_entries(C.byref(t), v2.ctypes.data)
    assert np.array_equal(v1, v2)
    t2 = T()
    r.copy_su2_tensor(C.byref(t), C.byref(t2))
    e.su2_tensor_deserialize_renormalized_entries(C.byref(t2), v2.ctypes.data)
    S.assert_same_su2(t2, t, 1e-15)


def _heisenberg_ground_state_sparse(L: int, J: float) -> float:
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    sx = sp.csr_matrix(np.array([[0, 0.5], [0.5, 0]])); sz = sp.csr_matrix(np.diag([0.5, -0.5]))
    sy = sp.csr_matrix(np.array([[0, -0.5j], [0.5j, 0]]))
    H = sp.csr_matrix((2 ** L, 2 ** L), dtype=complex)
    for i in range(L - 1):
        for s in (sx, sy, sz):
            H = H + J * sp.kron(sp.kron(sp.identity(2 ** i), sp.kron(s, s)), sp.identity(2 ** (L - i - 2)))
    return float(spl.eigsh(H.real.astype(float), k=1, which="SA")[0][0])


@pytest.mark.parametrize("kind", KINDS)
def test_su2_dmrg_beyond_the_reference_table_range(kind):
    """Bond quantum numbers above 2j = 5 (L = 14 chain, start bonds up to 2j = 5, no truncation): the reference reads past its recoupling
    tables there (src/tensor/su2_recoupling.c:959-963), so the check is the exact ground-state energy of the chain (sparse diagonalisation)."""
    e = S.engine(kind)
    L, ns = 14, 4
    mpo = S.heisenberg_mpo(L, 1.0)
    psi = S.random_mps(L, [1], [0, 1], 0, 5, 4, 77, scale=2.0)
    en = (C.c_double * ns)()
    ent = (C.c_double * (L - 1))()
    assert e.su2_dmrg_twosite(C.byref(mpo), ns, 12, 0.0, 1 << 20, C.byref(psi), en, ent) == 0
    top = max(psi.a[i].outer_irreps[2].jlist[k] for i in range(L) for k in range(psi.a[i].outer_irreps[2].num))
    assert top >= 6, top
    assert abs(en[ns - 1] - _heisenberg_ground_state_sparse(L, 1.0)) < 1e-9


@pytest.mark.parametrize("kind", KINDS)
def test_su2_mps_local_orthonormalize(kind):
    """one QR step to the right and one RQ step to the left: same bond structure as the reference, isometries, same two-site tensor"""
    r, e = S.ref(), S.engine(kind)
    r.su2_tensor_is_isometry.restype = C.c_bool
    r.su2_tensor_is_isometry.argtypes = [C.POINTER(T), C.c_double, C.c_bool]
    psi = S.random_mps(6, [1], [0, 1], 0, 5, 9, 23, scale=2.0)
    for rq in (False, True):
        i, j = (2, 3) if not rq else (3, 2)
        ar, br, ae, be = T(), T(), T(), T()
        for dst in (ar, ae):
            r.copy_su2_tensor(C.byref(psi.a[i]), C.byref(dst))
        for dst in (br, be):
            r.copy_su2_tensor(C.byref(psi.a[j]), C.byref(dst))
        if rq:
            r.su2_mps_local_orthonormalize_rq(C.byref(ar), C.byref(br)); e.su2_mps_local_orthonormalize_rq(C.byref(ae), C.byref(be))
        else:
            r.su2_mps_local_orthonormalize_qr(C.byref(ar), C.byref(br)); e.su2_mps_local_orthonormalize_qr(C.byref(ae), C.byref(be))
        S.assert_same_su2(ae, ar, 1e300)     # structure; the entries are unique up to the signs of the triangular factor
        S.assert_same_su2(be, br, 1e300)
        left, right = (ae, be) if not rq else (be, ae)
        pr, pe = T(), T()
        lr, rr = (ar, br) if not rq else (br, ar)
        r.su2_mps_contract_tensor_pair(C.byref(lr), C.byref(rr), C.byref(pr))
        r.su2_mps_contract_tensor_pair(C.byref(left), C.byref(right), C.byref(pe))
        S.assert_same_su2(pe, pr, 1e-12)


@pytest.mark.parametrize("kind", KINDS)
def test_su2_complex128(kind):
    """complex128 through the same kernels (conjugation of the bra tensor fused into the launches): a complex random SU(2) MPS against the
    Heisenberg MPO converted to complex entries, environments entry-wise and both DMRG variants against the reference"""
    r, e = S.ref(), S.engine(kind)
    L = 6
    mpo = S.heisenberg_mpo(L, 1.3)
    for i in range(L):
        S.complexify(mpo.a[i])
    psi = S.random_mps(L, [1], [0, 1], 0, 5, 7, 19, dtype=3, scale=3.0)
    rl_ref, rl_eng = (T * L)(), (T * L)()
    r.su2_compute_right_operator_blocks(C.byref(psi), C.byref(psi), C.byref(mpo), rl_ref)
    e.su2_compute_right_operator_blocks(C.byref(psi), C.byref(psi), C.byref(mpo), rl_eng)
    for i in range(L):
        S.assert_same_su2(rl_eng[i], rl_ref[i], 1e-12)
    lb = T()
    r.su2_create_dummy_operator_block_left(3, C.byref(lb))
    lbs = [lb]
    for i in range(L - 1):
        nr, ne = T(), T()
        r.su2_contraction_operator_step_left(C.byref(psi.a[i]), C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[-1]), C.byref(nr))
        e.su2_contraction_operator_step_left(C.byref(psi.a[i]), C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[-1]), C.byref(ne))
        S.assert_same_su2(ne, nr, 1e-12)
        lbs.append(nr)
    for i in range(L):
        br, be = T(), T()
        r.su2_apply_local_hamiltonian(C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[i]), C.byref(rl_ref[i]), C.byref(br))
        e.su2_apply_local_hamiltonian(C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[i]), C.byref(rl_ref[i]), C.byref(be))
        S.assert_same_su2(be, br, 1e-12)
    for two in (True, False):
        p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
        e1, e2 = (C.c_double * 2)(), (C.c_double * 2)()
        s1, s2 = (C.c_double * L)(), (C.c_double * L)()
        if two:
            assert r.su2_dmrg_twosite(C.byref(mpo), 2, 5, 1e-8, 100, C.byref(p1), e1, s1) == 0
            assert e.su2_dmrg_twosite(C.byref(mpo), 2, 5, 1e-8, 100, C.byref(p2), e2, s2) == 0
        else:
            assert r.su2_dmrg_singlesite(C.byref(mpo), 2, 5, C.byref(p1), e1) == 0
            assert e.su2_dmrg_singlesite(C.byref(mpo), 2, 5, C.byref(p2), e2) == 0
        assert np.allclose(list(e1), list(e2), rtol=0, atol=1e-11)
        assert r.su2_mps_is_consistent(C.byref(p2))


def test_su2_missing_charge_sectors():
    """A site tensor that lacks every charge sector of one bond quantum number (removed with the reference's
    su2_tensor_delete_charge_sector_by_index): orthonormalisation and both DMRG variants give the reference's norm, expectation value and energies.
    The engine drops the empty quantum number from the bond where the reference keeps it with an identity block: same state."""
    r, e = S.ref(), S.engine("emu")
    r.su2_tensor_delete_charge_sector_by_index.restype = None
    r.su2_tensor_delete_charge_sector_by_index.argtypes = [C.POINTER(T), C.c_int64]
    L = 6
    mpo = S.heisenberg_mpo(L, 1.0)
    psi = S.random_mps(L, [1], [0, 1], 0, 3, 4, 5, scale=3.0)
    secs = S.sectors(psi.a[2])
    assert any(s[2] == 3 for s in secs)
    for c in reversed(range(secs.shape[0])):
        if secs[c][2] == 3:
            r.su2_tensor_delete_charge_sector_by_index(C.byref(psi.a[2]), c)
    assert r.su2_mps_is_consistent(C.byref(psi))
    for mode in (0, 1):
        p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
        n1 = r.su2_mps_orthonormalize_qr(C.byref(p1), mode)
        n2 = e.su2_mps_orthonormalize_qr(C.byref(p2), mode)
        assert abs(n1 - n2) <= 1e-12 * n1 and r.su2_mps_is_consistent(C.byref(p2))
        a, b = C.c_double(), C.c_double()
        r.su2_mpo_inner_product(C.byref(p1), C.byref(mpo), C.byref(p1), C.byref(a))
        r.su2_mpo_inner_product(C.byref(p2), C.byref(mpo), C.byref(p2), C.byref(b))
        assert abs(a.value - b.value) <= 1e-12 * abs(a.value)
    for two in (True, False):
        p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
        e1, e2 = (C.c_double * 2)(), (C.c_double * 2)()
        s1, s2 = (C.c_double * L)(), (C.c_double * L)()
        if two:
            assert r.su2_dmrg_twosite(C.byref(mpo), 2, 5, 1e-8, 100, C.byref(p1), e1, s1) == 0
            assert e.su2_dmrg_twosite(C.byref(mpo), 2, 5, 1e-8, 100, C.byref(p2), e2, s2) == 0
        else:
            assert r.su2_dmrg_singlesite(C.byref(mpo), 2, 5, C.byref(p1), e1) == 0
            assert e.su2_dmrg_singlesite(C.byref(mpo), 2, 5, C.byref(p2), e2) == 0
        assert np.allclose(list(e1), list(e2), rtol=0, atol=1e-11)
        assert r.su2_mps_is_consistent(C.byref(p2))


def test_su2_and_u1_engines_agree_on_a_longer_chain():
    """Two independent code paths of the engine on the same physics at a size the SU(2) reference cannot run (L = 24, bonds up to 2j >= 6):
    the SU(2) two-site sweep on the Heisenberg chain against the U(1) two-site sweep on the XXZ chain with Delta = 1, h = 0, both from
    seeded random states with generous bonds; the converged energies agree to the truncation level."""
    e = S.engine("emu")
    eng = helpers.load("emu")
    ref = helpers.load("ref")
    L = 24
    mpo2 = S.heisenberg_mpo(L, 1.0)
    psi2 = S.random_mps(L, [1], [0, 1], 0, 5, 3, 11, scale=2.0)
    ns = 4
    en2 = (C.c_double * ns)()
    ent2 = (C.c_double * (L - 1))()
    assert e.su2_dmrg_twosite(C.byref(mpo2), ns, 8, 1e-10, 400, C.byref(psi2), en2, ent2) == 0
    top = max(psi2.a[i].outer_irreps[2].jlist[k] for i in range(L) for k in range(psi2.a[i].outer_irreps[2].num))
    assert top >= 4
    mpo1_r = helpers.ref_mpo(ref, "xxz", L, 1.0, 1.0, 0.0)
    psi1_r = helpers.ref_random_mps(ref, np.float64, L, mpo1_r.qsite, 0, 60, seed=7)
    mpo1, psi1 = helpers.clone_chain(eng, mpo1_r), helpers.clone_chain(eng, psi1_r)
    en1 = np.zeros(ns)
    ent1 = np.zeros(L - 1)
    assert eng.dmrg_twosite(mpo1.ptr, ns, 8, 1e-10, 120, psi1.ptr, en1.ctypes.data_as(C.POINTER(C.c_double)), ent1.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert abs(en2[ns - 1] - en1[ns - 1]) < 1e-7, (en2[ns - 1], en1[ns - 1])      # measured: 2e-9 (the U(1) state is the less converged one)
    # entanglement entropy of the centre bond: the SU(2) value counts every multiplet with its dimension
    assert abs(ent2[L // 2 - 1] - ent1[L // 2 - 1]) < 1e-5


@pytest.mark.parametrize("kind", KINDS)
def test_su2_dmrg_longer_chains_inside_the_reference_range(kind):
    """L = 32 single-site (start bonds up to 2j = 3) and L = 24 two-site (start bonds up to 2j = 2, tol_split 1e-5): the largest chains the
    unmodified reference runs without leaving its recoupling tables; energies against it"""
    r, e = S.ref(), S.engine(kind)
    mpo = S.heisenberg_mpo(32, 1.0)
    psi = S.random_mps(32, [1], [0, 1], 0, 3, 12, 42, scale=1.0)
    p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
    e1, e2 = (C.c_double * 2)(), (C.c_double * 2)()
    assert r.su2_dmrg_singlesite(C.byref(mpo), 2, 6, C.byref(p1), e1) == 0
    assert e.su2_dmrg_singlesite(C.byref(mpo), 2, 6, C.byref(p2), e2) == 0
    assert np.allclose(list(e1), list(e2), rtol=0, atol=1e-10)
    mpo = S.heisenberg_mpo(24, 1.0)
    psi = S.random_mps(24, [1], [0, 1], 0, 2, 3, 42, scale=1.0)
    p1, p2 = S.copy_mps(psi), S.copy_mps(psi)
    s1, s2 = (C.c_double * 23)(), (C.c_double * 23)()
    assert r.su2_dmrg_twosite(C.byref(mpo), 2, 6, 1e-5, 60, C.byref(p1), e1, s1) == 0
    assert e.su2_dmrg_twosite(C.byref(mpo), 2, 6, 1e-5, 60, C.byref(p2), e2, s2) == 0
    assert np.allclose(list(e1), list(e2), rtol=0, atol=1e-10)
    assert np.allclose(list(s1), list(s2), rtol=0, atol=1e-6)      # truncated at 1e-5: measured 1.5e-8


def _dense(r, t):
    """logical dense tensor (Clebsch-Gordan structure included) by the reference's su2_to_dense_tensor"""
    from chemtensor_b200 import cabi
    d = cabi.DenseTensor()
    r.su2_to_dense_tensor(C.byref(t), C.byref(d))
    shape = tuple(d.dim[i] for i in range(d.ndim))
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(C.cast(d.data, C.POINTER(C.c_double)), shape=(n,)).reshape(shape).copy()


@pytest.mark.parametrize("kind", KINDS)
def test_su2_local_hamiltonian_equals_the_dense_contraction(kind):
    """Pinned to the physics rather than to another SU(2) implementation (as the reference's test_su2_chain_ops.c does): the logical dense
    tensors of a, w, l, r contracted with einsum against the dense tensor of the engine's result, single-site and pair form"""
    r, e = S.ref(), S.engine(kind)
    e.ctb_su2_apply_local_hamiltonian_pair.restype = None
    e.ctb_su2_apply_local_hamiltonian_pair.argtypes = [C.POINTER(T)] * 6
    L = 5
    mpo = S.heisenberg_mpo(L, 1.3)
    psi = S.random_mps(L, [1], [0, 1], 1, 3, 3, 9, scale=2.0)
    lbs, rl = _envs(r, psi, mpo, L)
    for i in range(L):
        be = T()
        e.su2_apply_local_hamiltonian(C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[i]), C.byref(rl[i]), C.byref(be))
        a, w, l, rr, b = (_dense(r, x) for x in (psi.a[i], mpo.a[i], lbs[i], rl[i], be))
        want = np.einsum('olwk,lpr,wqps,rsmt->kqm', l, a, w, rr)
        assert np.max(np.abs(want - b)) <= 1e-13 * max(1.0, np.max(np.abs(want)))
    for i in range(L - 1):
        a2, b2 = T(), T()
        r.su2_mps_contract_tensor_pair(C.byref(psi.a[i]), C.byref(psi.a[i + 1]), C.byref(a2))
        e.ctb_su2_apply_local_hamiltonian_pair(C.byref(a2), C.byref(mpo.a[i]), C.byref(mpo.a[i + 1]), C.byref(lbs[i]), C.byref(rl[i + 1]), C.byref(b2))
        a, w0, w1, l, rr, b = (_dense(r, x) for x in (a2, mpo.a[i], mpo.a[i + 1], lbs[i], rl[i + 1], b2))
        want = np.einsum('olwk,lpqr,wxpu,uyqs,rsmt->kxym', l, a, w0, w1, rr)
        assert np.max(np.abs(want - b)) <= 1e-13 * max(1.0, np.max(np.abs(want)))


@pytest.mark.parametrize("kind", KINDS)
def test_su2_environment_steps_equal_the_dense_contraction(kind):
    """environment updates as dense einsum contractions of the logical tensors (real entries: no conjugation)"""
    r, e = S.ref(), S.engine(kind)
    L = 5
    mpo = S.heisenberg_mpo(L, 1.3)
    psi = S.random_mps(L, [1], [0, 1], 1, 3, 3, 9, scale=2.0)
    lbs, rl = _envs(r, psi, mpo, L)
    for i in range(L - 1, 0, -1):
        rn = T()
        e.su2_contraction_operator_step_right(C.byref(psi.a[i]), C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(rl[i]), C.byref(rn))
        a, w, rr, out = (_dense(r, x) for x in (psi.a[i], mpo.a[i], rl[i], rn))
        want = np.einsum('lpr,rsmt,wqps,kqm->lwkt', a, rr, w, a)
        assert np.max(np.abs(want - out)) <= 1e-13 * max(1.0, np.max(np.abs(want)))
    for i in range(L - 1):
        ln = T()
        e.su2_contraction_operator_step_left(C.byref(psi.a[i]), C.byref(psi.a[i]), C.byref(mpo.a[i]), C.byref(lbs[i]), C.byref(ln))
        a, w, l, out = (_dense(r, x) for x in (psi.a[i], mpo.a[i], lbs[i], ln))
        want = np.einsum('olwk,lpr,wqps,kqm->orsm', l, a, w, a)
        assert np.max(np.abs(want - out)) <= 1e-13 * max(1.0, np.max(np.abs(want)))
