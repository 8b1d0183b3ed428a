"""Drift check of the local Chebyshev scheme: chains run without any re-base (kpm_rebase_sweeps = 0) against chains that evaluate the full
trace for every proposal.  python tools/kpm_local_soak.py [chains] [sweeps]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import fk_mc_b200 as fk

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for kind, L, U, beta in (("cubic2d", 32, 2.0, 10.0), ("cubic2d", 16, 6.0, 4.0), ("triangular", 24, 1.0, 20.0)):
    res = {}
    for loc in (0, 1):
        c = fk.Context(kind, L, max_batch=chains)
        c.set_option("kpm_local", loc)
        c.set_option("kpm_rebase_sweeps", 0)
        c.chain_init(chains, beta, U, cheb_moves=True, seed=11, sweep_len=16, ntherm_sweeps=0, measure_energy=False, mc_flip=0.3, mc_add_remove=0.7,
                     record_trace=True, max_sweeps=sweeps)
        c.chain_run_sweeps(sweeps)
        res[loc] = (c.chain_get_trace(), c.chain_get_state())
        c.close()
    t0, t1 = res[0][0], res[1][0]
    same = np.array_equal(t0["accepted"], t1["accepted"]) and np.array_equal(res[0][1]["f"], res[1][1]["f"])
    rel = np.abs(t0["logz_new"] - t1["logz_new"]) / np.abs(t0["logz_new"])
    n = rel.shape[0]
    print(f"{kind} L={L} U={U} beta={beta}: {chains} chains x {sweeps * 16} proposals without re-base, {int(t1['accepted'].sum())} accepted moves: "
          f"chains identical {same}; logZ rel diff first quarter {rel[: n // 4].max():.2e}, last quarter {rel[-n // 4:].max():.2e}", flush=True)
