"""Worker of the world_size > 1 tests: one process per rank, gloo on CPU (emu engine) or NCCL on GPUs (cuda engine).

usage: python dist_worker.py <engine: emu|cuda> <outfile prefix>
Rank / world / rendezvous come from the torchrun-style environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT).
Every rank runs the same seeded inputs through the PUBLIC C API (apply_local_hamiltonian, dmrg_twosite) with the sharded
effective Hamiltonian switched on (ctb_dist_init) and stores its results; the parent test compares them with the
single-rank results and across ranks (they must be bit-identical: every rank holds the full state).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from chemtensor_b200 import cabi, workloads  # noqa: E402


def make_gloo_allgather(dist, torch, world):
    """all-gather callback for the CPU test double: raw host pointers -> numpy -> torch CPU tensors -> gloo"""
    def fn(ctx, sendbuf, recvbuf, nbytes, stream):
        send = np.ctypeslib.as_array(C.cast(sendbuf, C.POINTER(C.c_uint8)), shape=(nbytes,))
        recv = np.ctypeslib.as_array(C.cast(recvbuf, C.POINTER(C.c_uint8)), shape=(world * nbytes,))
        out = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(out, torch.from_numpy(send.copy()))
        for p in range(world):
            recv[p * nbytes:(p + 1) * nbytes] = out[p].numpy()
        return 0
    return cabi.ALLGATHER_FUNC(fn)


def main():
    kind, prefix = sys.argv[1], sys.argv[2]
    if len(sys.argv) > 3:
        mode = sys.argv[3]      # fused (NVSwitch multicast where available) | fused_uc (one store per peer) | pull | push | allgather
        if mode == "fused_uc":
            os.environ["CTB_NO_MULTICAST"] = "1"
            mode = "fused"
        os.environ["CTB_EXCHANGE"] = mode
        if mode in ("fused", "allgather"):
            os.environ["CTB_SHARDED_UPLOAD_MIN"] = "1"      # every host tensor goes up as 1/world per rank + all-gather
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if kind == "cuda":
        torch.cuda.set_device(rank)
        os.environ["CTB_DEVICE"] = str(rank)
    dist.init_process_group("gloo" if kind == "emu" else "nccl", rank=rank, world_size=world)
    eng = helpers.load(kind)
    keep = None
    if kind == "emu":
        assert eng.ctb_dist_init(rank, world, None) == 0
        keep = make_gloo_allgather(dist, torch, world)
        assert eng.ctb_dist_set_allgather(C.cast(keep, C.c_void_p), None) == 0
    else:
        # NCCL inside the C layer: rank 0 creates the unique id, the host distributes it
        uid = (C.c_uint8 * 128)()
        if rank == 0:
            assert eng.ctb_dist_unique_id(uid) == 0
        t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        uid = (C.c_uint8 * 128)(*t.cpu().tolist())
        assert eng.ctb_dist_init(rank, world, uid) == 0

    out = {}
    # 1. one sharded two-site Heff application on seeded operands (Fermi-Hubbard and XXZ structures, real and complex)
    for tag, model, L, params, sector, D, dtype in (("fh", "fermi_hubbard", 8, (1.0, 4.0, 0.0), workloads.encode_qpair(8, 0), 40, np.float64),
                                                     ("xxz", "xxz", 12, (1.0, 0.8, 0.1), 0, 24, np.complex128)):
        a, w, l, r, w0, w1 = workloads.heff_operands(eng, model, L, params, sector, D, dtype=dtype, seed=7, with_sites=True)
        b = cabi.BST(eng)
        eng.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr)
        out[f"heff_{tag}"] = b.serialize()
        # the same application in the pair form (two site tensors one after the other, no merged pair tensor), sharded as well
        b2 = cabi.BST(eng)
        assert eng.ctb_apply_local_hamiltonian_pair(a.ptr, w0.ptr, w1.ptr, l.ptr, r.ptr, b2.ptr) == 0
        out[f"heff_{tag}_pair"] = b2.serialize()
    # 2. a short two-site DMRG with every local solve sharded
    mpo = workloads.mpo_chain(eng, "fermi_hubbard", 6, (1.0, 4.0, 0.0))
    psi = workloads.random_mps(eng, np.float64, 6, mpo.qsite, workloads.encode_qpair(6, 0), 32, seed=42)
    en = np.zeros(2); ent = np.zeros(5)
    rc = eng.dmrg_twosite(mpo.ptr, 2, 12, 1e-10, 32, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    out["dmrg_en"] = en
    out["dmrg_entropy"] = ent
    # the same sweeps with every local solve in the pair form
    os.environ["CTB_HEFF_PAIR"] = "1"
    psi2 = workloads.random_mps(eng, np.float64, 6, mpo.qsite, workloads.encode_qpair(6, 0), 32, seed=42)
    en2 = np.zeros(2); ent2 = np.zeros(5)
    assert eng.dmrg_twosite(mpo.ptr, 2, 12, 1e-10, 32, psi2.ptr, en2.ctypes.data_as(C.POINTER(C.c_double)), ent2.ctypes.data_as(C.POINTER(C.c_double))) == 0
    del os.environ["CTB_HEFF_PAIR"]
    out["dmrg_en_pair"] = en2
    out["dmrg_site2"] = psi.site(2).serialize()
    info = (C.c_longlong * 4)()
    eng.ctb_dist_info(info)
    out["exchange_counts"] = np.array([info[2], info[3], eng.ctb_dist_pull_exchanges(), eng.ctb_dist_push_exchanges()], dtype=np.int64)      # fused peer-store, all-gather, pull, push exchanges
    out["multicast_exchanges"] = np.array([eng.ctb_dist_multicast_exchanges()], dtype=np.int64)
    np.savez(f"{prefix}_rank{rank}.npz", **out)
    eng.ctb_dist_finalize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
