import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, fk_mc_b200 as fk
B=1184
c=fk.Context("cubic2d",32,max_batch=B)
rng=np.random.default_rng(1)
f=(rng.random((B,1024))<0.5).astype(np.int32)
c.logz_kpm(f,2.0,1.0,20.0,16,32)
c.sync()
