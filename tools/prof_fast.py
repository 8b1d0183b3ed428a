"""A few fast-update steps of the c2 / c3 shape for ncu captures: python tools/prof_fast.py [workload] [chains]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fk_mc_b200 as fk
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, L, beta, U, chains = {"c2": ("cubic2d", 16, 10.0, 2.0, 4096), "c3": ("cubic3d", 8, 5.0, 4.0, 1024), "c4t": ("triangular", 24, 10.0, 2.0, 1024)}[wl]
if len(sys.argv) > 2:
    chains = int(sys.argv[2])
c = fk.Context(kind, L, max_batch=chains)
c.chain_init(chains, beta, U, seed=32167, sweep_len=4, ntherm_sweeps=100, measure_energy=False, max_sweeps=4, fast_update=True)
c.chain_run_sweeps(2)
print(c.chain_get_state()["naccept"].sum())
