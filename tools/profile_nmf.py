"""Small driver for ncu: the 160-problem rank sweep (k = 1..8 x 20 restarts) on 200 x 16 envelopes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import nmf_envelopes
from muscle_synergies_b200 import analysis

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
X = nmf_envelopes()
ranks = [k for k in range(1, 9) for _ in range(20)]
seeds = [r for _ in range(1, 9) for r in range(20)]
for _ in range(2):
    analysis.nmf_mu_batched(X, ranks, seeds, max_iter=iters, tol=0.0)
torch.cuda.synchronize()
print("done")
