"""Local KPM re-evaluation (kpm2d.cu) against the full trace: identical chains, logZ agreement, timing.
python tools/kpm_local_probe.py [chains] [sweeps]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import fk_mc_b200 as fk

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
for kind, L, flip in [("cubic2d", 32, 0.0), ("cubic2d", 32, 0.5), ("triangular", 24, 0.3), ("cubic2d", 16, 0.0)]:
    res = {}
    for loc in (0, 1):
        c = fk.Context(kind, L, max_batch=chains)
        c.set_option("kpm_local", loc)
        c.chain_init(chains, 10.0, 2.0, cheb_moves=True, seed=99, sweep_len=16, ntherm_sweeps=0, measure_energy=False, mc_flip=flip,
                     mc_add_remove=1.0 - flip, record_trace=True, max_sweeps=sweeps)
        c.chain_run_sweeps(sweeps)
        tr = c.chain_get_trace()
        st = c.chain_get_state()
        res[loc] = (tr, st)
        c.close()
    t0, t1 = res[0][0], res[1][0]
    same = np.array_equal(t0["accepted"], t1["accepted"]) and np.array_equal(res[0][1]["f"], res[1][1]["f"])
    rel = np.abs(t0["logz_new"] - t1["logz_new"]) / np.abs(t0["logz_new"])
    print(f"{kind} L={L} flip={flip}: chains identical {same}; logZ rel diff max {rel.max():.2e} mean {rel.mean():.2e}; accept rate {t0['accepted'].mean():.3f}", flush=True)
# timing at the headline shape
for loc in (0, 1):
    c = fk.Context("cubic2d", 32, max_batch=1024)
    c.set_option("kpm_local", loc)
    c.chain_init(1024, 10.0, 2.0, cheb_moves=True, seed=5, sweep_len=16, ntherm_sweeps=0, measure_energy=False, max_sweeps=40)
    c.chain_run_sweeps(2)
    c.profile_enable(True)
    c.profile_reset()
    t = time.time()
    c.chain_run_sweeps(8)
    dt = (time.time() - t) / 8
    print(f"kpm_local={loc}: {dt*1e3:.2f} ms per sweep of 16 KPM proposals x 1024 chains; moments {c.profile_get('kpm_moments')[0] / max(1, c.profile_get('kpm_moments')[1]):.3f} ms, "
          f"lanczos {c.profile_get('kpm_lanczos')[0] / max(1, c.profile_get('kpm_lanczos')[1]):.3f} ms per launch", flush=True)
    c.close()
