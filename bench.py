#!/usr/bin/env python
"""Headline benchmark: Metropolis proposals/s of the fk_mc weight-evaluation hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c5] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one sweep (sweep_len = 16 proposals per chain + the per-sweep measurement,
src/mc_metropolis.cpp:34-61) over every chain resident on the GPU.  Default workload = BASELINE's
headline configuration: cubic2d L=32 (N=1024), beta=20, U=2, Chebyshev/KPM moves (M=16, G=32) plus
one exact eigensolve per sweep for the energy measurement, 1024 independent chains per GPU.
`value` is device-timed with the chains resident in HBM; `e2e` drives the same sweep from HOST
buffers through the C-ABI evaluators (H2D of every proposal batch, D2H of every result).
Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (lattice, L, beta, U, cheb_moves, chains per GPU, description)
    "c5": ("cubic2d", 32, 20.0, 2.0, True, 1024, "cubic2d L=32 N=1024 beta=20 U=2, KPM moves M=16 G=32 + exact eigensolve per sweep"),
    "c2": ("cubic2d", 16, 10.0, 2.0, False, 4096, "cubic2d L=16 N=256 beta=10 U=2, dense eigensolve moves"),
    "c1": ("cubic2d", 8, 1.0, 1.0, False, 4096, "cubic2d L=8 N=64 beta=1 U=1, dense eigensolve moves"),
    "c3": ("cubic3d", 8, 5.0, 4.0, False, 1024, "cubic3d L=8 N=512 beta=5 U=4, dense eigensolve moves"),
    "c4t": ("triangular", 24, 10.0, 2.0, False, 1024, "triangular L=24 N=576 beta=10 U=2, dense eigensolve moves"),
    "c4h": ("honeycomb", 24, 10.0, 2.0, False, 1024, "honeycomb L=24 N=576 beta=10 U=2, dense eigensolve moves"),
}
SWEEP_LEN = 16      # src/mc_metropolis.cpp:14
SEED = 32167        # test/fast_update_test.cpp:48, benchmark/fast_update.cpp:152
CPU_SWEEPS_BY_WORKLOAD = {"c5": 8, "c1": 2000, "c2": 48, "c3": 8, "c4t": 6, "c4h": 6}  # sweeps per chain in one CPU sample: about 10 s per host thread
CPU_SWEEPS = 8
DMMA_PEAK_TFLOPS = 37.0  # measured on this pool's B200 by tools/probe_dmma.cu (profiles/r01_probe_dmma.txt)
DFMA_PEAK_TFLOPS = 36.8  # same probe, plain FP64 FMA


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                       str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def fk_cheb(L, ndim=2, prefactor=2.2):
    """fk_mc.hxx:60-63: M = even(int(ln N * prefactor)), G = max(2M, 10) (same as fk_mc_b200.cheb_sizes, without importing the package)."""
    m = int(math.log(float(L ** ndim)) * prefactor)
    m += m % 2
    return m, max(2 * m, 10)


def kpm_algorithmic_bytes(N, M):
    """streaming formulation, 24 N^2 (M/2 - 1) + 16 N^2 bytes per proposal."""
    return 24.0 * N * N * (M / 2 - 1) + 16.0 * N * N


def make_config(desc, chains, U, cheb, M, G):
    """The workload description both arms print (identical keys and values, so the driver can match them)."""
    return {"workload": desc, "chains_per_gpu": chains, "sweep_len": SWEEP_LEN, "seed": SEED, "mu_c": U / 2, "mu_f": U / 2,
            "moves": "add_remove", "M": M if cheb else None, "G": G if cheb else None,
            "l2_flush": "256 MiB device memset between timed steps", "parallelism": "chains sharded, %d per GPU" % chains}


def _lapack_worker(args):
    kind_id, L, U, budget_s, seed = args
    import scipy.linalg as sl
    from threadpoolctl import threadpool_limits
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as o
    n = o.lattice_size(kind_id, L)
    rng = np.random.default_rng(seed)
    H = o.hopping_dense(kind_id, L) + np.diag(U * (rng.random(n) < 0.5) - U / 2)
    with threadpool_limits(limits=1):
        sl.eigvalsh(H, driver="evd", check_finite=False)  # warm-up
        t0, k = time.perf_counter(), 0
        while time.perf_counter() - t0 < budget_s:
            sl.eigvalsh(H, driver="evd", check_finite=False)
            k += 1
        return k, time.perf_counter() - t0


def lapack_probe(kind, L, beta, U, cheb, cores, budget_s=6.0):
    """Runs in its own interpreter (no CUDA context, so forking workers is safe): `cores` processes, each solving with one thread."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as o
    kid = o.KINDS[kind]
    n = o.lattice_size(kid, L)
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_lapack_worker, [(kid, L, U, budget_s, i) for i in range(cores)])
    solves_per_s = sum(k / sec for k, sec in res)
    out = {"kind": "lapack dsyevd (scipy/OpenBLAS), one single-threaded process per core", "cores": cores, "solves_per_sec": solves_per_s,
           "sample": "%d solves of N=%d in %.1f s" % (sum(k for k, _ in res), n, max(sec for _, sec in res))}
    if cheb:
        M, G = fk_cheb(L, 3 if kind == "cubic3d" else (1 if kind == "cubic1d" else 2))
        f, _ = o.randomize_f(SEED, n, n // 2)
        t0 = time.perf_counter()
        reps = 4
        for _ in range(reps):
            o.calc_chebyshev(kid, L, f, U, U / 2, beta, M, G, emode=1)
        t_kpm = (time.perf_counter() - t0) / reps   # one thread; the port runs one chain per thread, so the rate scales with the cores
        sweep_s = SWEEP_LEN * t_kpm + cores / solves_per_s
        out["value"] = cores * SWEEP_LEN / sweep_s
        out["note"] = "16 port KPM evaluations (%.1f ms each) + one dsyevd (%.1f ms) per sweep and core" % (1e3 * t_kpm, 1e3 * cores / solves_per_s)
    else:
        out["value"] = solves_per_s
        out["note"] = "one dsyevd per proposal"
    out["unit"] = "proposals/s"
    return out


def lapack_baseline(workload, cores):
    """The stronger CPU reference of BASELINE.md section 3 / SURVEY 8d: LAPACK dsyevd (eigenvalues only) in place of the port's Eigen-style
    unblocked solver.  Reported beside the port; proposals/s the CPU arm would reach with it."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "lapack-probe", "--workload", workload, "--cores", str(cores)],
                             capture_output=True, text=True, timeout=300)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # the baseline is context, never a reason to lose the bench line
        return {"unavailable": repr(e)[:200]}


def run_reference(args, wl, rank, world):
    """CPU arm: the oracle restatement of the reference algorithm on all host cores (kind = "port":
    the reference cannot be built in this image).  P independent chains, chain r seeded SEED + r, as mpirun -np P would."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as o
    kind, L, beta, U, cheb, chains, desc = wl
    if args.chains:
        chains = args.chains
    ndim = 3 if kind == "cubic3d" else (1 if kind == "cubic1d" else 2)
    cores = os.cpu_count() or 1
    p = o.make_params(kind=o.KINDS[kind], L=L, beta=beta, U=U, cheb_moves=cheb, emode=1, seed=SEED, nsweeps=CPU_SWEEPS, sweep_len=SWEEP_LEN,
                      ntherm_sweeps=0, measure_energy=True)
    times = []
    for it in range(args.warmup + args.steps):
        sec, _ = o.bench_chains(p, cores, rank0=it * cores)
        if it >= args.warmup:
            times.append(sec)
    ms = 1e3 * float(np.mean(times))
    value = cores * SWEEP_LEN * CPU_SWEEPS / (ms * 1e-3)
    line = {"impl": "reference", "metric": "metropolis_proposals_per_sec", "value": value, "unit": "proposals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(desc, chains, U, cheb, *fk_cheb(L, ndim)),
            "cpu_baseline": {"value": value, "unit": "proposals/s", "cores": cores, "kind": "port",
                             "sample": "%d chains x %d sweeps (16 proposals + 1 measurement each) per step, one chain per host thread" % (cores, CPU_SWEEPS),
                             "lapack": lapack_baseline(args.workload, cores)},
            "e2e": {"value": value, "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "lapack-probe"])
    ap.add_argument("--cores", type=int, default=0, help="(lapack-probe) worker processes")
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--fast-update", action="store_true",
                    help="dense-move workloads: rank-one secular re-weighting with tracked eigenvectors (chain parameter fast_update)")
    ap.add_argument("--measure-ipr", action="store_true",
                    help="measurement path (c): every sweep also measures the IPR of all eigenstates (calc_ed(true) per chain, ipr.hpp:39-56)")
    ap.add_argument("--dense-path", action="store_true", help="full solves through the dense N^3 reduction (sy2sb) instead of the band path (sb2sb)")
    ap.add_argument("--kpm-full", action="store_true", help="Chebyshev moves evaluate the full trace for every proposal (option kpm_local = 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short dense-move lines (c2 / c3, full solve and fast update) appended to the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    global CPU_SWEEPS
    CPU_SWEEPS = CPU_SWEEPS_BY_WORKLOAD.get(args.workload, 8)
    kind, L, beta, U, cheb, chains, desc = wl
    if args.chains:
        chains = args.chains
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if args.impl == "lapack-probe":
        print(json.dumps(lapack_probe(kind, L, beta, U, cheb, args.cores or (os.cpu_count() or 1))), flush=True)
        return

    import torch
    import torch.distributed as dist
    import fk_mc_b200 as fk
    from fk_mc_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libfkmc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    stream = torch.cuda.Stream()
    ctx = fk.Context(kind, L, max_batch=chains, device=local_rank)
    if args.dense_path:
        ctx.set_option("band_path", 0)
    if args.kpm_full:
        ctx.set_option("kpm_local", 0)
    ctx.set_stream(stream.cuda_stream)
    N = ctx.N
    M, G = fk.cheb_sizes(N, 2.2)
    total_sweeps = args.warmup + args.steps
    chain0, _ = parallel.partition_chains(world * chains, world, rank)
    fast = bool(args.fast_update and not cheb)
    ctx.chain_init(chains, beta, U, cheb_moves=cheb, seed=SEED, chain0=chain0, sweep_len=SWEEP_LEN, ntherm_sweeps=0,
                   measure_energy=True, max_sweeps=total_sweeps, fast_update=fast, measure_ipr=args.measure_ipr,
                   measure_history=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            ctx.chain_run_sweeps(1)
            flush.zero_()
        stream.synchronize()
        ctx.profile_enable(True)
        ctx.profile_reset()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        launches0 = ctx.launch_count()
        nacc0 = int(ctx.chain_get_state()["naccept"].sum())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(args.steps):
            ctx.chain_run_sweeps(1)
            flush.zero_()  # L2 flush between timed steps (256 MiB write)
        ev1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        clocks = sampler.stop() if sampler else None
        launches = ctx.launch_count() - launches0 + args.steps  # + the flush kernels
        accepted = int(ctx.chain_get_state()["naccept"].sum()) - nacc0
        ctx.profile_enable(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    lt = torch.tensor([launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    proposals = world * chains * SWEEP_LEN * args.steps
    value = proposals / (ms_total * 1e-3)

    # per-kernel-family device time inside the timed region (events recorded on the launching stream)
    fam = {}
    for name in ("kpm", "sytrd", "sy2sb", "sb2sb", "band_build", "sb2st", "tridiag_eig", "build_h", "chain_step", "fu_eval", "fu_prepare", "fu_gemm", "fu_refresh", "stein",
                 "backtransform"):
        tot, n = ctx.profile_get(name)
        if n:
            fam[name] = {"ms_per_launch": tot / n, "launches": n, "share": tot / ms_total}
    sub = {}  # the two kernels inside one "kpm" evaluation (csrc/kpm2d.cu); their time is already counted in fam["kpm"]
    for name in ("kpm_lanczos", "kpm_moments"):
        tot, n = ctx.profile_get(name)
        if n:
            sub[name] = {"ms_per_launch": tot / n, "launches": n, "share": tot / ms_total, "part_of": "kpm"}
    peaks, peak_src = measured_peaks()
    traffic_file = None
    for cand in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        if os.path.exists(os.path.join(ROOT, "profiles", cand)):
            traffic_file = cand
            break
    # roofline of the dominant kernel (largest share of the timed region); both candidates are always reported
    rl_kpm = rl_dense = None
    if "kpm" in fam:
        bytes_launch = kpm_algorithmic_bytes(N, M) * chains
        achieved = bytes_launch / (fam["kpm"]["ms_per_launch"] * 1e-3) * 1e-9
        tr_k = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))["kpm"]
            tr_k = tr["bytes_per_launch"] * chains / tr["units_per_launch"]
        except Exception:
            pass
        # What binds the fused kernels is the FP64 pipe, not HBM: flops the algorithm needs WITH locality (T_m e_j lives on the sites within m hops
        # of j: 2m^2+2m+1 in 2-D, (2z+3) flop per element update, two patch dot products per column) plus the Lanczos steps actually taken
        zc = {"cubic2d": 4, "triangular": 6, "honeycomb": 3, "cubic3d": 6, "cubic1d": 2}[kind]
        half = M // 2
        if kind in ("cubic2d", "triangular", "honeycomb"):
            sup = lambda m: min(N, 2 * m * m + 2 * m + 1) if kind != "triangular" else min(N, 3 * m * m + 3 * m + 1)  # noqa: E731
        else:
            sup = lambda m: N  # noqa: E731  (3-D / 1-D run the full-lattice-vector kernel)
        f_mom = N * (sum(sup(m) * (2 * zc + 3) for m in range(2, half + 1)) + 4 * sup(half))
        try:
            lz_steps = float(np.mean(ctx.kpm_last_steps(chains)))
        except Exception:
            lz_steps = 135.0
        f_lz = 2 * lz_steps * N * (2 * zc + 7) / 2  # one Lanczos run delivers both e_min and e_max
        f_mom_full = f_mom
        kpm_local = kind in ("cubic2d", "triangular") and not getattr(args, "kpm_full", False)
        if kpm_local:
            # local scheme of the chain engine (csrc/kpm2d.cu): a one-site proposal recomputes the columns within M/2 - 1 hops of the changed site
            # for both configurations instead of all N columns
            n_aff = sup(half - 1)
            f_mom = 2 * n_aff * (sum(sup(m) * (2 * zc + 3) for m in range(2, half + 1)) + 4 * sup(half))
        a_fp = (f_mom + f_lz) * chains / (fam["kpm"]["ms_per_launch"] * 1e-3) * 1e-12
        rl_kpm = {"kernel": "lanczos2d_kernel + kpm_moments2d_kernel" if sub else "kpm_kernel", "bound": "fp64", "achieved": a_fp, "peak": DFMA_PEAK_TFLOPS,
                  "unit": "TFLOP/s", "frac": a_fp / DFMA_PEAK_TFLOPS, "traffic": tr_k, "peak_source": "measured FP64 DFMA (tools/probe_dmma.cu)",
                  "flop_per_proposal": {"moments": f_mom, "lanczos": f_lz, "lanczos_steps": lz_steps, "moments_full_trace": f_mom_full,
                                        "local_scheme": kpm_local},
                  "full_trace_equivalent_tflops": (f_mom_full + f_lz) * chains / (fam["kpm"]["ms_per_launch"] * 1e-3) * 1e-12,
                  "hbm_contract": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                                   "peak_source": peak_src,
                                   "note": "algorithmic bytes of the STREAMING formulation (SURVEY 8d: 24 N^2 (M/2-1) + 16 N^2 per proposal); the kernels keep "
                                           "the recursion in shared memory and touch ~8 KB of DRAM per proposal, so this ratio exceeds 1 and bounds nothing"},
                  "note": "the fused KPM kernels are bound by the FP64 pipe and shared-memory wavefronts (ncu: profiles/r01_ncu_summary_kpm_moments_v2.txt), "
                          "so the roofline is flops-with-locality against the measured DFMA peak"}
    dense_name = "sy2sb" if "sy2sb" in fam else ("sytrd" if "sytrd" in fam else None)
    if dense_name:
        fl = 4.0 / 3.0 * N ** 3 * chains
        a2 = fl / (fam[dense_name]["ms_per_launch"] * 1e-3) * 1e-12
        tr_d = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))[dense_name]
            tr_d = tr["bytes_per_launch"] * chains / tr["units_per_launch"]
        except Exception:
            pass
        rl_dense = {"kernel": dense_name + "_kernel", "bound": "tensor", "achieved": a2, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                    "frac": a2 / DMMA_PEAK_TFLOPS, "traffic": tr_d, "peak_source": "measured FP64 DMMA (tools/probe_dmma.cu)",
                    "note": "4/3 N^3 flop per matrix, all of it issued as DMMA.8x8x4 in the dense->band stage"}
    rl_band = None
    if "sb2sb" in fam:
        # band path (csrc/sb2sb.cu): the folded lattice matrix (half-bandwidth <= 64) to half-bandwidth 8 by block bulge chasing.  Useful flops of
        # one step (64 x 8 reflector block applied to a 64 x 56 block from the left, a 64 x 64 block from the right and a symmetric 64 x 64 block
        # from both sides, lower half): 4*64*8*56 + 4*64*8*64 + 2*64*64*8 + 2*64*65*8; the first step of a sweep has no left block
        nsteps = first = 0
        for j in range((N + 7) // 8):
            if 8 * j + 8 >= N:
                break
            k = (N - 8 * j - 8 + 63) // 64
            nsteps += k
            first += 1
        fl = (nsteps * (4 * 64 * 8 * 64 + 2 * 64 * 64 * 8 + 2 * 64 * 65 * 8) + (nsteps - first) * 4 * 64 * 8 * 56) * float(chains)
        a5 = fl / (fam["sb2sb"]["ms_per_launch"] * 1e-3) * 1e-12
        tr_b = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))["sb2sb"]
            tr_b = tr["bytes_per_launch"] * chains / tr["units_per_launch"]
        except Exception:
            pass
        rl_band = {"kernel": "sb2sb_kernel", "bound": "tensor", "achieved": a5, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": a5 / DMMA_PEAK_TFLOPS,
                   "traffic": tr_b, "peak_source": "measured FP64 DMMA (tools/probe_dmma.cu)", "steps_per_matrix": nsteps,
                   "dense_equivalent_tflops": 4.0 / 3.0 * N ** 3 * chains / (fam["sb2sb"]["ms_per_launch"] * 1e-3) * 1e-12,
                   "note": "band -> band reduction of the folded lattice matrix: %.3g flop per matrix instead of the 4/3 N^3 = %.3g of the dense reduction; "
                           "all of it DMMA.8x8x4; dense_equivalent_tflops = 4/3 N^3 over the same time" % (fl / chains, 4.0 / 3.0 * N ** 3)}
    rl_meas = None
    if args.measure_ipr and "sytrd" in fam:
        # one-stage tridiagonalisation that keeps its reflectors for the back-transformation: the SYMV streams the trailing matrix once per
        # column (8 N^3 / 6 bytes) and the rank-64 updates once per panel of 32 columns (read + write)
        per_launch = chains * args.steps / float(fam["sytrd"]["launches"])   # the eigenvector pipeline works through the batch in chunks
        by = 8.0 * N ** 3 / 6.0 * (1.0 + 2.0 / 32.0) * per_launch
        a4 = by / (fam["sytrd"]["ms_per_launch"] * 1e-3) * 1e-9
        rl_meas = {"kernel": "sytrd_lower_kernel", "bound": "hbm", "achieved": a4, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a4 / peaks["hbm_gbs"],
                   "traffic": None, "peak_source": peak_src,
                   "note": "measurement path (c): eigenvectors need the reflectors of the one-stage reduction, whose SYMV half is memory bound"}
    rl_fast = None
    if "fu_gemm" in fam:
        # eigenvector update of the accepted moves: 2 N^3 flop each, all DMMA (csrc/secular.cu: fu_gemm_kernel)
        tot_ms = fam["fu_gemm"]["ms_per_launch"] * fam["fu_gemm"]["launches"]
        a3 = 2.0 * N ** 3 * accepted / (tot_ms * 1e-3) * 1e-12
        rl_fast = {"kernel": "fu_gemm_kernel", "bound": "tensor", "achieved": a3, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                   "frac": a3 / DMMA_PEAK_TFLOPS, "traffic": None, "peak_source": "measured FP64 DMMA (tools/probe_dmma.cu)",
                   "note": "V <- V Q on accepted moves only: 2 N^3 flop x %d accepted of %d proposals in the timed region" % (accepted, chains * SWEEP_LEN * args.steps)}
    dominant = max(fam, key=lambda k: fam[k]["share"])
    roofline = rl_kpm if dominant == "kpm" else (rl_fast if (fast and rl_fast is not None) else (rl_dense if rl_dense is not None else rl_kpm))
    if dominant in ("sytrd", "backtransform", "stein") and rl_meas is not None:
        roofline = rl_meas
    if rl_band is not None and (dominant == "sb2sb" or (not fast and dominant not in ("kpm",) and rl_dense is None)):
        roofline = rl_band
    roofline_dense = rl_dense
    roofline_kpm = rl_kpm

    # ---- e2e: the same sweep driven from HOST buffers through the C-ABI evaluators ----
    e2e = None
    if not args.no_e2e:
        rng = np.random.default_rng(SEED + rank)
        f_pin = torch.zeros((chains, N), dtype=torch.int32).pin_memory()
        f_host = f_pin.numpy()
        f_host[:] = (rng.random((chains, N)) < 0.5)
        ebmu = math.exp(beta * U / 2)
        h2d = d2h = 0

        def pinned(*shape):
            return torch.zeros(shape, dtype=torch.float64).pin_memory().numpy()

        # result buffers: page-locked and reused from call to call (binding `out=`), as a host code would hold them
        out_kpm = dict(moments=pinned(chains, M), ab=pinned(chains, 4), logZ=pinned(chains), state=pinned(chains, 64)) if cheb else None
        out_ed = dict(spectrum=pinned(chains, N), logZ=pinned(chains))

        def weight_eval(fh, f_ref=None, state_ref=None):
            nonlocal h2d, d2h
            h2d += fh.nbytes
            if cheb:
                # fkmc_logz_kpm_batched_local: the proposal is evaluated against the configuration it was made from (the reference's
                # new_config = config; ...; new_config.calc_chebyshev()), whose trace-sum record travels with that configuration
                r = ctx.logz_kpm_local(fh, U, U / 2, beta, M, G, f_ref=f_ref, state_ref=state_ref, out=out_kpm)
                if f_ref is not None:
                    h2d += f_ref.nbytes + state_ref.nbytes
                d2h += r["moments"].nbytes + 4 * 8 * chains + r["logZ"].nbytes + r["state"].nbytes
            else:
                r = ctx.logz_ed(fh, U, U / 2, beta, out=out_ed)
                d2h += r["spectrum"].nbytes + r["logZ"].nbytes
            return r

        f_cur_pin = torch.zeros((chains, N), dtype=torch.int32).pin_memory()
        f_cur_host = f_cur_pin.numpy()
        f_cur_host[:] = f_host
        ks_cur = [None]

        def e2e_step(lz_cur):
            nonlocal h2d, d2h
            for _ in range(SWEEP_LEN):
                sites = rng.integers(0, N, size=chains)
                rows = np.arange(chains)
                f_host[rows, sites] ^= 1                      # propose in place
                r = weight_eval(f_host, f_cur_host if (cheb and ks_cur[0] is not None) else None, ks_cur[0] if cheb else None)
                lz_new = r["logZ"]
                occ = f_host[rows, sites] == 1
                w = np.exp(lz_new - lz_cur) * np.where(occ, ebmu, 1 / ebmu)
                acc = np.abs(w) > rng.random(chains)
                f_host[rows[~acc], sites[~acc]] ^= 1          # reject: undo
                f_cur_host[rows[acc], sites[acc]] ^= 1        # accept: the current configuration (and its record) follows
                if cheb:
                    ks_cur[0][acc] = r["state"][acc]
                lz_cur = np.where(acc, lz_new, lz_cur)
            if args.measure_ipr:                              # measurement sweep with eigenvectors: spectrum + IPR of every eigenstate
                h2d += f_host.nbytes
                r = ctx.ipr(f_host, U, U / 2, beta)
                d2h += r["spectrum"].nbytes + r["ipr"].nbytes
            elif cheb:                                        # measurement sweep: exact spectrum -> energy
                h2d += f_host.nbytes
                r = ctx.logz_ed(f_host, U, U / 2, beta, out=out_ed)
                d2h += r["spectrum"].nbytes + r["logZ"].nbytes
            return lz_cur

        r0 = weight_eval(f_host)
        lz = r0["logZ"].copy()
        if cheb:
            ks_cur[0] = pinned(chains, 64)
            ks_cur[0][:] = r0["state"]
        lz = e2e_step(lz)  # warm-up
        h2d = d2h = 0
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            lz = e2e_step(lz)
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3
        te = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": proposals / (float(te.item()) * 1e-3), "unit": "proposals/s", "h2d_bytes_per_step": h2d // args.steps,
               "d2h_bytes_per_step": d2h // args.steps,
               "path": "fkmc_logz_kpm_batched_local / fkmc_logz_ed_batched with pinned host f and pinned, reused result buffers; host-side proposal + accept"}

    # ---- final collective: gather the per-chain series (the only inter-GPU traffic of a run) ----
    gather_ms = None
    if world > 1:
        # fkmc_gather_series: ncclAllGather straight from the chain engine's device buffers (the library resolves libnccl itself);
        # torch.distributed only carries the 128-byte communicator id between the processes
        ids = [fk.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], world, rank)
        ctx.gather_series()  # warm-up (communicator set-up)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        allv = ctx.gather_series()
        gather_ms = (time.perf_counter() - t0) * 1e3
        assert allv["energies"].shape == (total_sweeps, world * chains)
        mine = ctx.chain_get_series()["energies"]
        assert np.array_equal(allv["energies"][:, rank * chains:(rank + 1) * chains], mine)

    # ---- CPU baseline on the box's host cores (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as o
        cores = os.cpu_count() or 1
        p = o.make_params(kind=o.KINDS[kind], L=L, beta=beta, U=U, cheb_moves=cheb, emode=1, seed=SEED, nsweeps=CPU_SWEEPS,
                          sweep_len=SWEEP_LEN, ntherm_sweeps=0, measure_energy=True)
        sec, _ = o.bench_chains(p, cores)
        cpu_baseline = {"value": cores * SWEEP_LEN * CPU_SWEEPS / sec, "unit": "proposals/s", "cores": cores, "kind": "port",
                        "sample": "%d chains x %d sweeps (16 proposals + 1 measurement each), one chain per host thread, %.1f s"
                                  % (cores, CPU_SWEEPS, sec),
                        "lapack": lapack_baseline(args.workload, cores)}

    # ---- the dense-move configurations of BASELINE.json next to the headline (rank 0, N = 1, default workload only): each is this script
    #      run on that workload for a few sweeps, so that the driver's JSON carries measured numbers for them as well ----
    extra = None
    if rank == 0 and world == 1 and args.workload == "c5" and not args.no_extra and not args.measure_ipr:
        extra = {}
        for wl_name, fastflag in (("c2", False), ("c2", True), ("c3", False), ("c3", True)):
            cmd = [sys.executable, os.path.abspath(__file__), "--workload", wl_name, "--no-cpu-baseline", "--no-e2e", "--no-extra", "--steps", "3", "--warmup", "3"]
            if fastflag:
                cmd.append("--fast-update")
            key = wl_name + ("_fast_update" if fastflag else "_full_solve")
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
                d = json.loads(r.stdout.strip().splitlines()[-1])
                extra[key] = {"workload": d["config"]["workload"], "chains_per_gpu": d["config"]["chains_per_gpu"], "value": d["value"], "unit": d["unit"],
                              "ms_per_step": d["ms_per_step"], "accept_rate": d.get("accept_rate"), "dominant_kernel": d["dominant_kernel"],
                              "roofline": {k: d["roofline"].get(k) for k in ("kernel", "bound", "achieved", "peak", "unit", "frac")},
                              "kernels": {k: {"ms_per_launch": v["ms_per_launch"], "share": v["share"]} for k, v in d["kernels"].items()}}
            except Exception as e:  # never lose the headline line over an extra
                extra[key] = {"unavailable": repr(e)[:200]}

    if rank == 0:
        line = {"metric": "metropolis_proposals_per_sec", "value": value, "unit": "proposals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(make_config(desc, chains, U, cheb, M, G), **({"fast_update": True} if fast else {}), **({"dense_path": True} if args.dense_path else {}),
                               **({"measure_ipr": True} if args.measure_ipr else {})),
                "accept_rate": accepted / float(chains * SWEEP_LEN * args.steps),
                "sweeps_per_sec": value / SWEEP_LEN, "roofline": roofline, "roofline_dense": roofline_dense, "roofline_kpm": roofline_kpm,
                "roofline_fast_update": rl_fast, "roofline_measurement": rl_meas, "roofline_band": rl_band, "traffic_source": traffic_file,
                "dominant_kernel": dominant, "kernels": {**fam, **sub},
                "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(lt.item()), "clocks": clocks,
                "final_gather_ms": gather_ms, "dense_move_workloads": extra}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
