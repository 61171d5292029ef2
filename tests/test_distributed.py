"""World_size > 1 path: the effective Hamiltonian sharded over the bra bond of the right environment,
one all-gather per application.  CPU: two gloo ranks driving the host logic on the test double; GPU (-m gpu, needs >= 2
devices): two NCCL ranks driving the CUDA product.  Results must equal the single-rank results to rounding (the shards
only change which GPU computes a column, not the order of any summation) and be bit-identical across ranks."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi, workloads

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_world(kind, world, tmp_path, mode="fused"):
    port = free_port()
    prefix = str(tmp_path / "dist")
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), kind, prefix, mode], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o.decode(errors="replace"))
    for rank, p in enumerate(procs):
        assert p.returncode == 0, f"rank {rank} failed:\n{outs[rank][-3000:]}"
    return [np.load(f"{prefix}_rank{r}.npz") for r in range(world)]


def single_rank_results(eng):
    out = {}
    for tag, model, L, params, sector, D, dtype in (("fh", "fermi_hubbard", 8, (1.0, 4.0, 0.0), workloads.encode_qpair(8, 0), 40, np.float64),
                                                     ("xxz", "xxz", 12, (1.0, 0.8, 0.1), 0, 24, np.complex128)):
        a, w, l, r = workloads.heff_operands(eng, model, L, params, sector, D, dtype=dtype, seed=7)
        b = cabi.BST(eng)
        eng.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr)
        out[f"heff_{tag}"] = b.serialize()
    mpo = workloads.mpo_chain(eng, "fermi_hubbard", 6, (1.0, 4.0, 0.0))
    psi = workloads.random_mps(eng, np.float64, 6, mpo.qsite, workloads.encode_qpair(6, 0), 32, seed=42)
    en = np.zeros(2); ent = np.zeros(5)
    assert eng.dmrg_twosite(mpo.ptr, 2, 12, 1e-10, 32, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double))) == 0
    out["dmrg_en"] = en
    return out


def check(results, single):
    for key in ("heff_fh", "heff_xxz"):
        for r in results:
            assert helpers.rel_err(r[key], single[key]) <= 1e-13
        for r in results[1:]:
            assert np.array_equal(r[key], results[0][key])      # every rank holds the same full vector, bit for bit
    for r in results:
        # pair form (no merged MPO pair tensor), sharded: same vector as the merged form on one rank
        for key in ("heff_fh", "heff_xxz"):
            assert helpers.rel_err(r[key + "_pair"], single[key]) <= 1e-12
        assert np.max(np.abs(r["dmrg_en_pair"] - single["dmrg_en"])) <= 1e-10
    for r in results:
        assert np.max(np.abs(r["dmrg_en"] - single["dmrg_en"])) <= 1e-10
    for r in results[1:]:
        assert np.array_equal(r["dmrg_en"], results[0]["dmrg_en"])
        assert np.array_equal(r["dmrg_site2"], results[0]["dmrg_site2"])


def assert_exchange_mode(r, mode):
    counts = tuple(int(x) > 0 for x in r["exchange_counts"])      # fused, all-gather, pull, push
    want = {"fused": (True, False, False, False), "fused_uc": (True, False, False, False), "allgather": (False, True, False, False), "pull": (False, False, True, False), "push": (False, False, False, True)}[mode]
    assert counts == want, (mode, r["exchange_counts"])
    if mode == "fused_uc":
        assert int(r["multicast_exchanges"][0]) == 0


@pytest.mark.parametrize("world,mode", [(2, "fused"), (3, "fused"), (2, "fused_uc"), (2, "allgather"), (2, "pull"), (3, "pull"), (2, "push"), (3, "push")])
def test_sharded_heff_and_dmrg_gloo(tmp_path, world, mode):
    """fused: step 3 stores into the result buffers of all ranks -- through one multicast store per element (NVSwitch multicast on GPUs,
    its shared-memory double here) or, fused_uc, one store per peer-mapped buffer; allgather: all-gather of the slices + scatter"""
    results = run_world("emu", world, tmp_path, mode)
    check(results, single_rank_results(helpers.load("emu")))
    for r in results:
        assert_exchange_mode(r, mode)
        if mode == "fused":
            assert int(r["multicast_exchanges"][0]) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "fused_uc", "allgather", "pull", "push"])
def test_sharded_heff_and_dmrg_nccl(tmp_path, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    results = run_world("cuda", 2, tmp_path, mode)
    check(results, single_rank_results(helpers.load("cuda")))
    for r in results:
        assert_exchange_mode(r, mode)
    print(f"mode {mode}: multicast exchanges per rank {[int(r['multicast_exchanges'][0]) for r in results]}")
