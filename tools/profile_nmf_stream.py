"""Small driver for ncu: the long-signal (streaming) NMF regime, 2 M x 16, k = 8, 4 problems."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muscle_synergies_b200 import analysis
n, m, k, P = 2_000_000, 16, 8, 4
rng = np.random.default_rng(0)
X = torch.rand((n, m), device="cuda", dtype=torch.float32)
init = [(np.abs(rng.standard_normal((n, k))).astype(np.float32), np.abs(rng.standard_normal((k, m))).astype(np.float32)) for _ in range(P)]
analysis.nmf_mu_batched(X, [k] * P, list(range(P)), max_iter=6, tol=0.0, init=init, regime="stream")
torch.cuda.synchronize()
print("done")
