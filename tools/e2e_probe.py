"""Where does the time of bench.py's host-driven (e2e) sweep go?  c5 shape: cubic2d L = 32, KPM moves M = 16 / G = 32, 1024 chains.
Modes: base (fresh pageable result arrays per call, as bench.py did), pinned (reused page-locked result buffers), split G (G contexts on G
streams driven by G host threads, chains divided between them; ctypes releases the GIL during the calls).
usage: python tools/e2e_probe.py [sweeps]"""
import math
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fk_mc_b200 as fk  # noqa: E402

L, N, U, beta, M, G, SWEEP_LEN = 32, 1024, 2.0, 20.0, 16, 32, 16
CH = 1024
sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 6


def pinned(shape, dtype=torch.float64):
    return torch.zeros(shape, dtype=dtype).pin_memory().numpy()


class Driver:
    def __init__(self, chains, seed, use_pinned, device=0):
        self.ctx = fk.Context("cubic2d", L, max_batch=chains, device=device)
        self.chains = chains
        self.rng = np.random.default_rng(seed)
        self.f = pinned((chains, N), torch.int32)
        self.f[:] = self.rng.random((chains, N)) < 0.5
        self.fc = pinned((chains, N), torch.int32)
        self.fc[:] = self.f
        self.out = self.out_ed = None
        if use_pinned:
            self.out = dict(moments=pinned((chains, M)), ab=pinned((chains, 4)), logZ=pinned((chains,)), state=pinned((chains, 64)))
            self.out_ed = dict(spectrum=pinned((chains, N)), logZ=pinned((chains,)))
            self.ks = pinned((chains, 64))
        else:
            self.ks = np.zeros((chains, 64))
        r = self.ctx.logz_kpm_local(self.f, U, U / 2, beta, M, G, out=self.out)
        self.lz = r["logZ"].copy()
        self.ks[:] = r["state"]
        self.t = dict(host=0.0, kpm=0.0, ed=0.0)

    def sweep(self):
        rows = np.arange(self.chains)
        ebmu = math.exp(beta * U / 2)
        for _ in range(SWEEP_LEN):
            t0 = time.perf_counter()
            sites = self.rng.integers(0, N, size=self.chains)
            self.f[rows, sites] ^= 1
            t1 = time.perf_counter()
            r = self.ctx.logz_kpm_local(self.f, U, U / 2, beta, M, G, f_ref=self.fc, state_ref=self.ks, out=self.out)
            t2 = time.perf_counter()
            lz_new = r["logZ"]
            occ = self.f[rows, sites] == 1
            w = np.exp(lz_new - self.lz) * np.where(occ, ebmu, 1 / ebmu)
            acc = np.abs(w) > self.rng.random(self.chains)
            self.f[rows[~acc], sites[~acc]] ^= 1
            self.fc[rows[acc], sites[acc]] ^= 1
            self.ks[acc] = r["state"][acc]
            self.lz = np.where(acc, lz_new, self.lz)
            t3 = time.perf_counter()
            self.t["host"] += (t1 - t0) + (t3 - t2)
            self.t["kpm"] += t2 - t1
        t0 = time.perf_counter()
        self.ctx.logz_ed(self.f, U, U / 2, beta, out=self.out_ed)
        self.t["ed"] += time.perf_counter() - t0


def run(label, groups, use_pinned):
    drv = [Driver(CH // groups, 1234 + g, use_pinned) for g in range(groups)]
    for d in drv:
        d.sweep()
        d.t = dict(host=0.0, kpm=0.0, ed=0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if groups == 1:
        for _ in range(sweeps):
            drv[0].sweep()
    else:
        def work(d):
            for _ in range(sweeps):
                d.sweep()
        th = [threading.Thread(target=work, args=(d,)) for d in drv]
        for t in th:
            t.start()
        for t in th:
            t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ms = dt / sweeps * 1e3
    parts = {k: round(sum(d.t[k] for d in drv) / groups / sweeps * 1e3, 2) for k in ("host", "kpm", "ed")}
    print("%-22s %7.2f ms per sweep  %8.1f k proposals/s   per-thread ms: %s" % (label, ms, CH * SWEEP_LEN / ms, parts), flush=True)
    for d in drv:
        d.ctx.close()


if __name__ == "__main__":
    torch.cuda.init()
    run("base", 1, False)
    run("pinned", 1, True)
    run("pinned split 2", 2, True)
    run("pinned split 4", 4, True)
