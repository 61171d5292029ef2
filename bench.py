#!/usr/bin/env python
"""bench.py -- Heff matvec FP64 TFLOP/s (and two-site DMRG sweep seconds) of the B200 engine vs the reference's CPU path.

Contract (one JSON line on stdout, rank 0):
  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME] [--sweep]

A "step" is one application of the block-sparse two-site effective Hamiltonian (reference apply_local_hamiltonian,
src/algorithm/chain_ops.c:353) at the centre bond of the named synthetic Hamiltonian:
  value  : algorithmic FP64 flops of one matvec (sum 2 m n k over the block GEMMs the reference issues at
           block_sparse_tensor.c:1953-1994; x4 for complex128) / device time per matvec, operands resident in HBM,
           cached per-bond plans, CUDA events on the engine's stream, L2 flushed between timed matvecs.
  e2e    : the same metric through the reference-named C-ABI call apply_local_hamiltonian() on HOST structs:
           host->device copies of (a, w, l, r), plan construction, three grouped-GEMM launches, device->host copy of b.
  --impl reference : the unmodified reference (oracle/_ref, built by oracle/Makefile) running the same call on the
           box's host cores (OpenMP over blocks, all cores).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from chemtensor_b200 import cabi, workloads  # noqa: E402

# measured bond structures: name -> (data file, L of the measured chain, scale of the multiplicities)
MEASURED = {
    "fh_L64_D4096": ("bonds_fh_L32_D2048.json", 32, 2),
    "fh_L32_D1024": ("bonds_fh_L32_D1024.json", 32, 1),
    "xxz_L100_D1024": ("bonds_xxz_L100_D1024.json", 100, 1),
}

WORKLOADS = {
    # name: (model, nsites, params, sector, max_vdim, dtype, description)
    "xxz_L100_D1024": ("xxz", 100, (1.0, 0.8, 0.1), 0, 1024, np.float64,
                       "Heisenberg XXZ chain L=100 (J=1, D=0.8, h=0.1), U(1) 2Sz=0, bond dim 1024, two-site Heff at the centre bond"),
    "fh_L64_D4096": ("fermi_hubbard", 64, (1.0, 4.0, 0.0), workloads.encode_qpair(64, 0), 4096, np.float64,
                     "Fermi-Hubbard chain L=64 (t=1, u=4, mu=0), U(1)xU(1) N=64 2Sz=0, bond dim 4096, two-site Heff at the centre bond"),
    "fh_L32_D1024": ("fermi_hubbard", 32, (1.0, 4.0, 0.0), workloads.encode_qpair(32, 0), 1024, np.float64,
                     "Fermi-Hubbard chain L=32, N=32 2Sz=0, bond dim 1024, two-site Heff at the centre bond"),
    "xxz_L16_D64": ("xxz", 16, (1.0, 0.8, 0.1), 0, 64, np.float64, "small XXZ case for functional checks"),
}
DEFAULT_WORKLOAD = "fh_L64_D4096"

CUDA_SO = os.path.join(ROOT, "chemtensor_b200", "libchemtensor_b200.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libchemtensor_ref.so")


def load_measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, device: int):
        self.device = device
        self.samples = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = max(float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples)}


def measure_fp64_peak(device: int):
    """cuBLAS DGEMM 8192^3 through torch: the FP64 tensor-pipe roofline denominator (MEASURED_PEAKS.json has no FP64 entry)."""
    import torch
    n = 8192
    with torch.cuda.device(device):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        del a, b
        torch.cuda.empty_cache()
    return best


def build_operands(lib, wl_name: str, seed: int = 42, structure: str = "converged"):
    """(a, w, l, r) of the centre pair.  structure = "converged": the sector histogram a converged two-site sweep of this
    model produced (measured, chemtensor_b200/data), multiplicities scaled to the named bond dimension;
    "random": the structure of the reference's construct_random_mps (uniformly sub-sampled sectors, many tiny blocks)."""
    model, L, params, sector, D, dtype, _ = WORKLOADS[wl_name]
    bonds = None
    if structure == "converged" and wl_name in MEASURED:
        data, L0, scale = MEASURED[wl_name]
        site0 = L0 // 2 - 1
        dN = (L - L0) // 2      # same filling: the particle number left of the centre grows with half the extra sites
        shift = workloads.encode_qpair(dN, 0) if model == "fermi_hubbard" else 0
        bonds = (workloads.measured_bond_qnums(data, site0, scale, shift, D), workloads.measured_bond_qnums(data, site0 + 2, scale, shift, D))
    return workloads.heff_operands(lib, model, L, params, sector, D, dtype=dtype, seed=seed, bonds=bonds)


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def run_reference(args):
    """The reference's own CPU implementation of the same call, on all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    if not os.path.exists(REF_SO):
        subprocess.run(["make", "-s", "-C", ROOT, "oracle"], check=False)
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the reference must get all host cores, so the variable is
    # set BEFORE the library (and with it libgomp) is loaded, and the count the runtime really uses is what gets reported
    want = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(want)
    os.environ.pop("OMP_PROC_BIND", None)
    ref = cabi.CLibrary(REF_SO)
    cores = omp_threads(want)
    a, w, l, r = build_operands(ref, wl, structure=args.structure)
    flops = heff_flops_host(a, w, l, r)
    # bound the CPU work: cap the number of timed calls so that the arm ends within a few minutes
    t_probe0 = time.perf_counter()
    b = cabi.BST(ref)
    ref.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr)
    t_probe = time.perf_counter() - t_probe0
    del b
    budget_s = 120.0
    steps = max(1, min(args.steps, int(budget_s / max(t_probe, 1e-6))))
    warm = 0 if t_probe > 20 else min(args.warmup, 1)
    for _ in range(warm):
        b = cabi.BST(ref); ref.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr); del b
    t0 = time.perf_counter()
    for _ in range(steps):
        b = cabi.BST(ref); ref.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr); del b
    dt = (time.perf_counter() - t0) / steps
    tf = flops / dt / 1e12
    line = {
        "impl": "reference", "metric": "heff_matvec_fp64_tflops", "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "description": WORKLOADS[wl][6], "structure": args.structure},
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "reference",
                         "sample": f"{steps} timed apply_local_hamiltonian calls of the unmodified reference (OpenMP {cores} threads, OpenBLAS 1 thread per GEMM) on the same operands"},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def omp_threads(want: int) -> int:
    """Set and read back the OpenMP thread count of this process (libgomp is already mapped by the reference library)."""
    try:
        gomp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
        gomp.omp_set_num_threads(C.c_int(want))
        gomp.omp_get_max_threads.restype = C.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return int(os.environ.get("OMP_NUM_THREADS", "1"))


def heff_flops_host(a, w, l, r) -> float:
    """Algorithmic flops of one matvec from sector metadata alone (integer exact): the three contractions of chain_ops.c:353-390."""
    from chemtensor_b200.flops import heff_flops
    return heff_flops(a, w, l, r)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CTB_BENCH_WORKLOAD", DEFAULT_WORKLOAD), choices=sorted(WORKLOADS))
    ap.add_argument("--structure", default="converged", choices=["converged", "random"],
                    help="sector structure of the synthetic operands: measured from converged sweeps (default) or the random-MPS rule")
    ap.add_argument("--sweep", dest="sweep", action="store_true", default=True, help="report two-site DMRG sweep seconds as well (default; single GPU)")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="skip the sweep block (profiling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    rank, world, local = dist_setup(args.gpus)
    if world > 1:
        # a rank that is terminated from outside (launcher timeout) says where it was
        import faulthandler, signal
        faulthandler.register(signal.SIGTERM, all_threads=True, chain=True)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    os.environ["CTB_DEVICE"] = str(local)
    lib = cabi.CLibrary(CUDA_SO, extensions=True)
    if lib.ctb_init(local) < 0:
        raise SystemExit("bench.py: ctb_init failed")
    if world > 1:
        # one process per GPU: NCCL communicator inside the C layer; rank 0 creates the unique id, torch.distributed carries it
        import torch.distributed as dist
        uid = (C.c_uint8 * 128)()
        if rank == 0 and lib.ctb_dist_unique_id(uid) < 0:
            raise SystemExit("bench.py: ctb_dist_unique_id failed")
        t_uid = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
        dist.broadcast(t_uid, 0)
        uid = (C.c_uint8 * 128)(*t_uid.cpu().tolist())
        if lib.ctb_dist_init(rank, world, uid) < 0:
            raise SystemExit("bench.py: ctb_dist_init failed")
    wl = args.workload
    model, L, params, sector, D, dtype, desc = WORKLOADS[wl]

    t_setup0 = time.perf_counter()
    a, w, l, r = build_operands(lib, wl, structure=args.structure)
    n_vec = a.num_elements()
    t_setup = time.perf_counter() - t_setup0

    peaks, peak_kind = load_measured_peaks()
    fp64_peak = measure_fp64_peak(local) if rank == 0 else 0.0

    launches0 = lib.ctb_launch_count()
    ms = C.c_double(0); flops = C.c_double(0)
    step_ms = (C.c_double * 3)(); step_fl = (C.c_double * 3)()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); torch.cuda.synchronize()
    with ClockSampler(local) as clocks:
        rc = lib.ctb_heff_benchmark(a.ptr, w.ptr, l.ptr, r.ptr, args.warmup, args.steps, 1, C.byref(ms), C.byref(flops), step_ms, step_fl)
        if rc < 0:
            raise SystemExit("bench.py: ctb_heff_benchmark failed")
        if world > 1:
            torch.cuda.synchronize(); dist.barrier()
        # whole-job algorithmic flops: the shards partition the block GEMMs column-wise, so their flops add up exactly
        flops_total = flops.value
        if world > 1:
            tf_ = torch.tensor([flops.value], dtype=torch.float64, device="cuda")
            dist.all_reduce(tf_, op=dist.ReduceOp.SUM)
            flops_total = float(tf_.item())
        # e2e through the reference-named C-ABI entry point with host structs
        e2e = None
        b_sharded = None
        if world > 1 and rank == 0 and not args.no_cpu_baseline:
            b = cabi.BST(lib); lib.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr); b_sharded = b.serialize(); del b
        elif world > 1:
            b = cabi.BST(lib); lib.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr); del b      # the call is collective
        if not args.no_e2e:
            esize = np.dtype(dtype).itemsize
            h2d = sum(x.num_elements() for x in (a, w, l, r)) * esize
            d2h = n_vec * esize
            for _ in range(2):
                b = cabi.BST(lib); lib.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr); del b
            k_e2e = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                b = cabi.BST(lib); lib.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr); del b
            dt_e2e = (time.perf_counter() - t0) / k_e2e
            if world > 1:
                te = torch.tensor([dt_e2e], dtype=torch.float64, device="cuda")
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
                dt_e2e = float(te.item())
            e2e = {"value": flops_total / dt_e2e / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": dt_e2e * 1e3, "steps": k_e2e,
                   "what": "apply_local_hamiltonian(a, w, l, r, &b) on host structs: upload, per-bond plan build, 3 grouped GEMM launches, download"}
    launches = lib.ctb_launch_count() - launches0

    # strong scaling: ONE matvec sharded over the ranks (bra bond of the right environment cut into tile-aligned, work-balanced index
    # sets; per application one exchange of the result slices over NVLink peer memory); time = max over ranks of the device time,
    # value = whole-job flops / that time
    t_ms = ms.value
    if world > 1:
        tt = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt.item())
    value = flops_total / (t_ms * 1e-3) / 1e12

    # strong scaling of whole two-site sweeps: every rank runs the same dmrg_twosite call (sharded matvec, sector-sharded SVD)
    sweep_dist = None
    if args.sweep and world > 1:
        sweep_dist = []
        for name, model, L_, params_, sector_, D_, nsw, _ in SWEEP_CASES[:2]:
            nsw = 1 if D_ >= 4096 else nsw
            rec = {"config": name, "sweeps": nsw, "lanczos_iterations": 10, "tol_split": 0.0, "n_gpus": world,
                   "sharded": "effective-Hamiltonian applications by bra-bond column slices (exchange as in config.multi_gpu); SVD split by sector blocks dealt to the ranks by cost, one all-gather of the factors; environments, plans and level-1 work replicated"}
            if os.environ.get("CTB_BENCH_VERBOSE"):
                print(f"[bench rank {rank}] sweep case {name} starts", file=sys.stderr, flush=True)
            rec["b200"] = sweep_seconds(lib, model, L_, params_, sector_, D_, sweeps=nsw)
            if os.environ.get("CTB_BENCH_VERBOSE"):
                print(f"[bench rank {rank}] sweep case {name} done: {rec['b200'] and rec['b200']['s_per_sweep']}", file=sys.stderr, flush=True)
            sweep_dist.append(rec)
    if rank != 0:
        return
    kdom = int(np.argmax([step_ms[i] for i in range(3)]))
    k_tf = step_fl[kdom] / (step_ms[kdom] * 1e-3) / 1e12
    names = ["A.R (step 1)", "W.(AR) (step 2)", "L.(WAR) (step 3)"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if world == 1 and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("workload") == wl and tj.get("structure") == args.structure:
            kk = tj["kernels"][["step1_grouped_gemm", "step2_mix_kernel", "step3_grouped_gemm"][kdom]]
            traffic = kk["dram_read_bytes"] + kk["dram_write_bytes"]      # bytes per launch of the dominant kernel, from the committed ncu capture
    roofline = {"bound": "tensor", "achieved": k_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (k_tf / fp64_peak) if fp64_peak > 0 else None, "traffic": traffic,
                "kernel": f"grouped_gemm_kernel<double> of {names[kdom]}", "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "per_step_ms": [step_ms[i] for i in range(3)], "per_step_tflops": [step_fl[i] / (step_ms[i] * 1e-3) / 1e12 if step_ms[i] > 0 else None for i in range(3)],
                "matvec_frac_of_fp64_peak": (flops_total / world / (t_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak > 0 else None,
                "exchange_ms": (ms.value - sum(step_ms[i] for i in range(3))) if world > 1 else 0.0}
    info = (C.c_double * 8)()
    if lib.ctb_heff_plan_info(a.ptr, w.ptr, l.ptr, r.ptr, rank, world, info) == 0 and step_ms[1] > 0:
        esize_ = np.dtype(dtype).itemsize
        mix_bytes = (info[6] + info[7]) * esize_      # one read of t1 + one write of t2 (algorithmic bytes of the MPO-mixing launch)
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline["secondary"] = {"kernel": "mix_kernel<double> of W.(AR) (step 2)", "bound": "hbm", "achieved": mix_bytes / (step_ms[1] * 1e-3) / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": mix_bytes / (step_ms[1] * 1e-3) / 1e9 / hbm_peak, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                                 "algorithmic_bytes": mix_bytes}
    line = {
        "metric": "heff_matvec_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if dtype == np.float64 else "c128",
        "data": "synthetic",
        "config": {"workload": wl, "description": desc, "structure": args.structure, "vector_length": int(n_vec), "flops_per_matvec": flops_total,
                   "l2": "192 MiB buffer rewritten between timed matvecs", "setup_s": round(t_setup, 2),
                   "multi_gpu": (f"one matvec sharded over {world} GPUs by tile-aligned, work-balanced index sets of the bra bond of the right environment; "
                                 f"exchange of the result slices per application: {exchange_name(lib)}") if world > 1 else "single GPU"},
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if args.sweep and world == 1:
        line["sweep"] = sweep_report(lib)
    elif sweep_dist is not None:
        line["sweep"] = sweep_dist
    if not args.no_cpu_baseline:
        base = cpu_baseline(wl, flops_total, args.structure, dump=b_dump_path(wl) if world > 1 else None)
        line["cpu_baseline"] = base
        if world > 1 and b_sharded is not None and base.get("value") is not None:
            # the sharded result of this very run against the unmodified reference on the same operands (north_star: 1e-12 relative)
            b_ref = np.load(b_dump_path(wl))
            err = float(np.linalg.norm(b_sharded - b_ref) / np.linalg.norm(b_ref))
            line["parity_checked"] = bool(err <= 1e-12)
            line["parity_rel_err_vs_reference"] = err
            os.remove(b_dump_path(wl))
            if not line["parity_checked"]:
                print(json.dumps(line), flush=True)
                raise SystemExit(f"bench.py: sharded result differs from the reference by {err:.3e}")
    print(json.dumps(line), flush=True)


def exchange_name(lib) -> str:
    info = (C.c_longlong * 4)()
    lib.ctb_dist_info(info)
    if info[2] > 0 and lib.ctb_dist_multicast_exchanges() > 0:
        return "fused over NVSwitch multicast (step-3 GEMM epilogue stores every element ONCE to the multicast address with multimem.st, the switch replicates it into the result buffers of all GPUs; peer-flag barrier)"
    if info[2] > 0:
        return "fused (step-3 GEMM epilogue stores into the peer-mapped result buffers of all GPUs, peer-flag barrier)"
    if lib.ctb_dist_push_exchanges() > 0:
        return "push (local step 3, one copy kernel stores the slice into the peer-mapped result buffers, peer-flag barrier)"
    if lib.ctb_dist_pull_exchanges() > 0:
        return "pull (peer-mapped send buffers read over NVLink after the barrier)"
    return "NCCL all-gather + scatter kernel"


def sweep_seconds(lib, model, L, params, sector, D, sweeps=2, lanczos=10, tol=0.0):
    """Wall seconds per two-site DMRG sweep (the other half of BASELINE.json's metric) through the public dmrg_twosite() on host structs,
    from the seeded random MPS; also returns the energies so that both arms can be compared."""
    mpo = workloads.mpo_chain(lib, model, L, params)
    psi = workloads.random_mps(lib, np.float64, L, mpo.qsite, sector, D, seed=42)
    en = np.zeros(sweeps); ent = np.zeros(L - 1)
    t0 = time.perf_counter()
    rc = lib.dmrg_twosite(mpo.ptr, sweeps, lanczos, tol, D, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
    dt = time.perf_counter() - t0
    if rc != 0:
        return None
    out = {"s_per_sweep": dt / sweeps, "energies": [float(x) for x in en], "max_bond": int(max(psi.bond_dims()))}
    ph = phases(lib)
    if ph is not None:
        out["phases_s"] = ph
    return out


def phases(lib):
    """Host wall-clock per phase of the last dmrg_* call of the engine (each phase ends with a device sync)."""
    if not lib.has("ctb_get_stats"):
        return None
    st = (C.c_double * 19)()
    lib.ctb_get_stats(st, 19)
    return {"per_sweep_s": [st[11 + i] / 1e3 for i in range(8) if st[11 + i] > 0], "lanczos_incl_plans": st[3] / 1e3, "svd_or_qr_split": st[4] / 1e3, "environments": st[5] / 1e3, "total": st[6] / 1e3,
            "host_plan_building_within_phases": st[9] / 1e3, "host_reblocking_calls_within_phases": st[10] / 1e3,
            "heff_calls": int(st[1]), "heff_tflop": st[0] / 1e12, "max_vector_len": int(st[7])}


SWEEP_CASES = [
    # (name, model, L, params, sector, D, sweeps, time the reference too?)
    ("fh_L64_D4096", "fermi_hubbard", 64, (1.0, 4.0, 0.0), workloads.encode_qpair(64, 0), 4096, 2, False),      # BASELINE.json configs[2], the north-star target
    ("fh_L32_D1024", "fermi_hubbard", 32, (1.0, 4.0, 0.0), workloads.encode_qpair(32, 0), 1024, 2, False),
    ("fh_L16_D256", "fermi_hubbard", 16, (1.0, 4.0, 0.0), workloads.encode_qpair(16, 0), 256, 2, True),
]


def molecular_sweep_seconds(lib, n, D, pair_form, sweeps=2, lanczos=10, seed=5):
    """BASELINE.json configs[3] in small: complex128 two-site DMRG on a synthetic molecular Hamiltonian (random complex integrals with
    the Hermitian symmetrisation of the reference's perf/perf_dmrg_coeffs.py, spatial orbitals, d = 4).  The MPO comes from the
    reference's own generator (oracle/_ref, input generator only).  pair_form: apply the two site MPO tensors one after the other
    (no merged pair tensor) -- the form large MPO bonds need; None leaves the engine's automatic choice."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    ref = helpers.load("ref")
    rng = np.random.default_rng(seed)
    tkin = 0.5 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    vint = 0.1 * (rng.standard_normal((n, n, n, n)) + 1j * rng.standard_normal((n, n, n, n)))
    tkin = 0.5 * (tkin + tkin.conj().T)
    vint = 0.5 * (vint + vint.transpose((1, 0, 3, 2)))
    vint = 0.5 * (vint + vint.transpose((2, 3, 0, 1)).conj())
    mpo_r = helpers.ref_molecular_mpo(ref, tkin, vint, spin=True, optimize=False)
    L = mpo_r.nsites
    psi_r = helpers.ref_random_mps(ref, np.complex128, L, mpo_r.qsite, workloads.encode_qpair(n, 0), D, seed=42)
    mpo, psi = helpers.clone_chain(lib, mpo_r), helpers.clone_chain(lib, psi_r)
    if pair_form is not None:
        os.environ["CTB_HEFF_PAIR"] = "1" if pair_form else "0"
    en = np.zeros(sweeps); ent = np.zeros(L - 1)
    t0 = time.perf_counter()
    rc = lib.dmrg_twosite(mpo.ptr, sweeps, lanczos, 0.0, D, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
    dt = time.perf_counter() - t0
    os.environ.pop("CTB_HEFF_PAIR", None)
    if rc != 0:
        return None
    return {"s_per_sweep": dt / sweeps, "energies": [float(x) for x in en], "max_bond": int(max(psi.bond_dims())), "mpo_bond": int(max(mpo_r.bond_dims()))}


def perf_dmrg_c1(lib):
    """BASELINE.json configs[0] = the reference's perf/perf_dmrg.c as shipped (:45-104): spin molecular Hamiltonian of 9 orbitals from
    the integrals of perf/perf_dmrg_coeffs.py (bit-identical under numpy's default_rng(42)), d = 4, sector (N = 9, 2Sz = 1),
    seed_rng_state(42) random MPS with max_vdim = 512, dmrg_twosite with 2 sweeps, 25 Lanczos iterations, tol_split = 1e-8.  The MPO
    and the start state come from the reference's own generators (oracle/_ref, input generation only); the reference's energy on
    every thread count is -51.2777797066802."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    ref = helpers.load("ref")
    tkin, vint = helpers.perf_dmrg_coeffs()
    mpo_r = helpers.ref_molecular_mpo(ref, tkin, vint, spin=True, optimize=False)
    L = mpo_r.nsites
    psi_r = helpers.ref_random_mps(ref, np.float64, L, mpo_r.qsite, workloads.encode_qpair(L, 1), 512, seed=42)
    mpo, psi = (mpo_r, psi_r) if lib is ref else (helpers.clone_chain(lib, mpo_r), helpers.clone_chain(lib, psi_r))
    en = np.zeros(2); ent = np.zeros(L - 1)
    t0 = time.perf_counter()
    rc = lib.dmrg_twosite(mpo.ptr, 2, 25, 1e-8, 512, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
    dt = time.perf_counter() - t0
    if rc != 0:
        return None
    out = {"wall_s": dt, "energies": [float(x) for x in en], "bond_dims": [int(x) for x in psi.bond_dims()], "mpo_bond_dims": [int(x) for x in mpo_r.bond_dims()]}
    ph = phases(lib) if lib is not ref else None
    if ph is not None:
        out["phases_s"] = ph
    return out


def xxz_c2(lib, single_site: bool, sweeps: int = 1, lanczos: int = 10):
    """BASELINE.json configs[1]: Heisenberg XXZ chain L = 100 (J = 1, D = 0.8, h = 0.1), U(1) sector 2Sz = 0, max bond 1024,
    tol_split = 0 (bonds saturate), from the seeded random MPS; two-site sweep or single-site sweep (dmrg_singlesite)."""
    model, L, params, sector, D = "xxz", 100, (1.0, 0.8, 0.1), 0, 1024
    if not single_site:
        return sweep_seconds(lib, model, L, params, sector, D, sweeps=sweeps, lanczos=lanczos)
    mpo = workloads.mpo_chain(lib, model, L, params)
    psi = workloads.random_mps(lib, np.float64, L, mpo.qsite, sector, D, seed=42)
    en = np.zeros(sweeps)
    t0 = time.perf_counter()
    rc = lib.dmrg_singlesite(mpo.ptr, sweeps, lanczos, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)))
    dt = time.perf_counter() - t0
    if rc != 0:
        return None
    out = {"s_per_sweep": dt / sweeps, "energies": [float(x) for x in en], "max_bond": int(max(psi.bond_dims()))}
    ph = phases(lib)
    if ph is not None:
        out["phases_s"] = ph
    return out


MOLECULAR_LARGE_CASES = [
    # (name, spatial orbitals, bond dimension)
    ("mol_n24_D1024_c128", 24, 1024),
]
# (name, arguments of tools/su2_run.py, time the reference too)
SU2_CASES = [
    ("su2_heisenberg_L24_parity", "24 60 --sweeps 2 --lanczos 6 --degen 3 --max-irrep 2 --tol 1e-5", True),
    # three growth sweeps by default (about 18 s: the bond reaches ~1200 multiplets / logical 6144); CTB_BENCH_SU2_FULL=1 adds the fourth,
    # saturated sweep (~2000 multiplets, logical 8192, about 70 s more) -- recorded in profiles/r2_bench_1gpu_su2_sweep_block.json
    ("su2_heisenberg_L200_2048multiplets", "200 8192 --sweeps %d --lanczos 10 --degen 8" % (4 if os.environ.get("CTB_BENCH_SU2_FULL") else 3), False),
]

MOLECULAR_SWEEP_CASES = [
    # (name, spatial orbitals, bond dimension, time the reference too?)
    ("mol_n8_D64_c128", 8, 64, True),
    ("mol_n12_D256_c128", 12, 256, False),
]


def _sweep_selected(name: str) -> bool:
    """CTB_BENCH_SWEEP_ONLY=substr[,substr...] restricts the sweep block to the matching configs (profiling aid; default: all)."""
    only = os.environ.get("CTB_BENCH_SWEEP_ONLY")
    return only is None or any(tok and tok in name for tok in only.split(","))


def sweep_report(lib):
    """Two-site sweep seconds of the engine (and of the unmodified reference on the small case) -- reported, not the headline."""
    out = []
    for name, model, L, params, sector, D, nsw, with_ref in SWEEP_CASES:
        if not _sweep_selected(name):
            continue
        rec = {"config": name, "sweeps": nsw, "lanczos_iterations": 10, "tol_split": 0.0, "start": "seeded random MPS (construct_random_mps rule), bonds saturate at max_vdim"}
        ours = sweep_seconds(lib, model, L, params, sector, D, sweeps=nsw)
        rec["b200"] = ours
        if with_ref and os.path.exists(REF_SO):
            code = (
                "import sys, json\n"
                f"sys.path.insert(0, {ROOT!r})\n"
                "import bench\n"
                "from chemtensor_b200 import cabi\n"
                f"ref = cabi.CLibrary({REF_SO!r})\n"
                f"print(json.dumps(bench.sweep_seconds(ref, {model!r}, {L}, {params!r}, {sector}, {D}, sweeps={nsw})))\n"
            )
            env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
            try:
                r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
                rec["reference_cpu"] = json.loads(r.stdout.strip().splitlines()[-1])
                rec["reference_cpu"]["cores"] = os.cpu_count()
                if ours is not None:
                    rec["max_energy_diff_vs_reference"] = float(np.max(np.abs(np.array(ours["energies"]) - np.array(rec["reference_cpu"]["energies"]))))
                    rec["energy_parity_1e-10"] = bool(rec["max_energy_diff_vs_reference"] <= 1e-10)
            except Exception as exc:
                rec["reference_cpu"] = {"failed": str(exc)}
        out.append(rec)
    # BASELINE.json configs[0]: perf/perf_dmrg.c as shipped, engine and reference side by side
    if _sweep_selected("perf_dmrg_c1"):
        rec = {"config": "perf_dmrg_c1", "what": "perf/perf_dmrg.c as shipped: 9 orbitals, sector (9, 1), max_vdim 512, 2 sweeps x 25 Lanczos iterations, tol_split 1e-8",
               "expected_energy": -51.2777797066802}
        rec["b200"] = perf_dmrg_c1(lib)
        if rec["b200"] is not None:
            rec["b200_repeat_wall_s"] = (perf_dmrg_c1(lib) or {}).get("wall_s")      # second call: plans' device allocations come from the warm pool
            rec["energy_error_vs_expected"] = abs(rec["b200"]["energies"][-1] - rec["expected_energy"])
        if os.path.exists(REF_SO):
            code = (
                "import sys, json\n"
                f"sys.path.insert(0, {ROOT!r})\n"
                f"sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
                "import bench, helpers\n"
                "print(json.dumps(bench.perf_dmrg_c1(helpers.load('ref'))))\n"
            )
            env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
            try:
                r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
                rec["reference_cpu"] = json.loads(r.stdout.strip().splitlines()[-1])
                rec["reference_cpu"]["cores"] = os.cpu_count()
                if rec["b200"] is not None:
                    rec["max_energy_diff_vs_reference"] = float(np.max(np.abs(np.array(rec["b200"]["energies"]) - np.array(rec["reference_cpu"]["energies"]))))
                    rec["energy_parity_1e-10"] = bool(rec["max_energy_diff_vs_reference"] <= 1e-10)
            except Exception as exc:
                rec["reference_cpu"] = {"failed": str(exc)}
        out.append(rec)
    # BASELINE.json configs[1]: XXZ L=100, D=1024, two-site and single-site
    if _sweep_selected("xxz_L100_D1024_twosite"):
        out.append({"config": "xxz_L100_D1024_twosite", "sweeps": 1, "lanczos_iterations": 10, "tol_split": 0.0, "b200": xxz_c2(lib, False)})
    if _sweep_selected("xxz_L100_D1024_singlesite"):
        out.append({"config": "xxz_L100_D1024_singlesite", "sweeps": 1, "lanczos_iterations": 10, "b200": xxz_c2(lib, True)})
    for name, n, D, with_ref in MOLECULAR_SWEEP_CASES:
        if not _sweep_selected(name):
            continue
        rec = {"config": name, "sweeps": 2, "lanczos_iterations": 10, "tol_split": 0.0, "dtype": "c128"}
        rec["b200_merged_pair_tensor"] = molecular_sweep_seconds(lib, n, D, False)
        rec["b200_pair_form"] = molecular_sweep_seconds(lib, n, D, True)
        if with_ref and os.path.exists(REF_SO):
            code = (
                "import sys, json\n"
                f"sys.path.insert(0, {ROOT!r})\n"
                "import bench\n"
                "from chemtensor_b200 import cabi\n"
                f"ref = cabi.CLibrary({REF_SO!r})\n"
                f"print(json.dumps(bench.molecular_sweep_seconds(ref, {n}, {D}, None)))\n"
            )
            env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
            try:
                r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
                rec["reference_cpu"] = json.loads(r.stdout.strip().splitlines()[-1])
                rec["reference_cpu"]["cores"] = os.cpu_count()
            except Exception as exc:
                rec["reference_cpu"] = {"failed": str(exc)}
        out.append(rec)
    # BASELINE.json configs[3] at the largest size this round runs: 24 spatial orbitals (MPO bond 2 n^2 + 3 n + 2 = 1226), complex128,
    # D = 1024, ONE sweep from the seeded random MPS.  The merged pair tensor of the reference would have 1226 x 256 x 1226 entries
    # per bond; the engine switches to the pair form (two site MPO tensors applied one after the other) by itself.
    for name, n, D in MOLECULAR_LARGE_CASES:
        if not _sweep_selected(name):
            continue
        rec = {"config": name, "sweeps": 1, "lanczos_iterations": 10, "tol_split": 0.0, "dtype": "c128", "form": "pair form chosen by the engine (merged pair tensor > 2^28 entries)"}
        try:
            rec["b200"] = molecular_sweep_seconds(lib, n, D, None, sweeps=1)
            if rec["b200"] is not None:
                rec["b200"]["phases_s"] = phases(lib)
        except Exception as exc:
            rec["b200"] = {"failed": str(exc)}
        out.append(rec)
    # BASELINE.json configs[4]: SU(2)-symmetric Heisenberg chain, SU(2) DMRG variant (su2_dmrg_twosite behind the reference's symbol).
    # (i) a chain the unmodified reference can run (its recoupling tables end at 2j = 5, src/tensor/su2_recoupling.c:959-963: start bonds
    # up to 2j = 2, tol_split 1e-5) for energy parity and the CPU time beside it; (ii) L = 200 as named: logical bond dimension 8192 = about 2000 multiplets once the bond
    # has saturated in the fourth sweep (the first three sweeps grow it from 8 multiplets per sector).
    for name, su2_args, with_ref in SU2_CASES:
        if not _sweep_selected(name):
            continue
        rec = {"config": name, "args": su2_args, "driver": "tools/su2_run.py (one su2_dmrg_twosite call on host structs; max_vdim is the logical bond dimension)"}
        outp = f"/tmp/ctb_bench_su2_{name}_{os.getpid()}.json"
        cmd = [sys.executable, os.path.join(ROOT, "tools", "su2_run.py"), os.environ.get("CTB_BENCH_SU2_ENGINE", "cuda")] + su2_args.split() + ["--out", outp] + (["--ref"] if with_ref else [])
        try:
            env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
            subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
            src = outp if os.path.exists(outp) else outp + ".partial"     # a reference that dies leaves the engine's sweeps in the partial record
            with open(src) as f:
                got = json.load(f)
            rec["b200"] = got.get("engine")
            if with_ref:
                rec["reference_cpu"] = got.get("reference", {"failed": "the reference did not finish"})
                if "energy_diff" in got:
                    rec["max_energy_diff_vs_reference"] = got["energy_diff"]
                    rec["energy_parity_1e-10"] = bool(got["energy_diff"] <= 1e-10)
        except Exception as exc:
            rec["b200"] = {"failed": str(exc)}
        out.append(rec)
    return out


def b_dump_path(wl: str) -> str:
    return f"/dev/shm/ctb_bench_ref_b_{wl}_{os.getpid()}.npy"


def cpu_baseline(wl: str, flops: float, structure: str = "converged", dump=None):
    """The unmodified reference (oracle/_ref) on the host cores of this box, bounded sample, rank 0 only.  dump: file the reference's
    result vector goes to (parity check of the sharded runs)."""
    cores = os.cpu_count() or 1
    code = (
        "import os,sys,time,json\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "from chemtensor_b200 import cabi\n"
        f"ref = cabi.CLibrary({REF_SO!r})\n"
        f"a,w,l,r = bench.build_operands(ref, {wl!r}, structure={structure!r})\n"
        "ts=[]\n"
        "t_all=time.perf_counter()\n"
        "while len(ts) < 5 and time.perf_counter()-t_all < 25:\n"
        "    t0=time.perf_counter(); b=cabi.BST(ref); ref.apply_local_hamiltonian(a.ptr,w.ptr,l.ptr,r.ptr,b.ptr); ts.append(time.perf_counter()-t0); del b\n"
        f"dump = {dump!r}\n"
        "if dump:\n"
        "    import numpy as np\n"
        "    b=cabi.BST(ref); ref.apply_local_hamiltonian(a.ptr,w.ptr,l.ptr,r.ptr,b.ptr); np.save(dump, b.serialize())\n"
        "gomp = __import__('ctypes').CDLL('libgomp.so.1'); gomp.omp_get_max_threads.restype = __import__('ctypes').c_int\n"
        "print(json.dumps({'ts': ts, 'threads': int(gomp.omp_get_max_threads())}))\n"
    )
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_PROC_BIND"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=1200)
        rec = json.loads(out.stdout.strip().splitlines()[-1])
        ts = rec["ts"]
        cores = int(rec.get("threads", cores))
        best = min(ts)
        return {"value": flops / best / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "reference", "ms_per_step": best * 1e3,
                "sample": f"best of {len(ts)} apply_local_hamiltonian calls of the unmodified reference (oracle/_ref, OpenMP {cores} threads) on the same operands"}
    except Exception as exc:  # the baseline is a reported number, never a gate
        return {"value": None, "unit": "TFLOP/s", "cores": cores, "kind": "reference", "sample": f"failed: {exc}"}


if __name__ == "__main__":
    main()
