"""Two half-batches of chains on two streams, out of phase, against one batch on one stream (c5 configuration).
python tools/overlap_probe.py [total_chains] [sweeps]"""
import sys, time, threading
import numpy as np
sys.path.insert(0, ".")
import torch
import fk_mc_b200 as fk

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
L, beta, U = 32, 10.0, 1.0

def make(chains, chain0):
    s = torch.cuda.Stream()
    c = fk.Context("cubic2d", L, max_batch=chains)
    c.set_stream(s.cuda_stream)
    c.chain_init(chains, beta, U, cheb_moves=True, seed=1234, chain0=chain0, sweep_len=16, ntherm_sweeps=0, measure_energy=True,
                 max_sweeps=sweeps + 3, measure_history=False)
    return c, s

def run(groups, offset_phase):
    # warm up
    for c, s in groups:
        c.chain_run_sweeps(1)
    torch.cuda.synchronize()
    def worker(c, n):
        c.chain_run_sweeps(n)
    t0 = time.time()
    th = [threading.Thread(target=worker, args=(c, sweeps)) for c, s in groups]
    for i, t in enumerate(th):
        t.start()
        if offset_phase and i + 1 < len(th):
            time.sleep(offset_phase)
    for t in th:
        t.join()
    torch.cuda.synchronize()
    return (time.time() - t0) / sweeps * 1e3

one = [make(total, 0)]
ms1 = run(one, 0)
e1 = one[0][0].chain_get_series()["energies"]
one[0][0].close()
print(f"1 group  x {total}: {ms1:.1f} ms per sweep", flush=True)
for ng in (2, 4):
    g = [make(total // ng, i * (total // ng)) for i in range(ng)]
    ms = run(g, 0.02)
    e = np.concatenate([c.chain_get_series()["energies"] for c, s in g], axis=1)
    same = np.array_equal(e[: e1.shape[0]], e1[: e.shape[0]])
    print(f"{ng} groups x {total // ng}: {ms:.1f} ms per sweep (whole job), series identical to the single group: {same}", flush=True)
    for c, s in g:
        c.close()
