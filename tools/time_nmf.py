"""Times the resident NMF kernel for several builds of the library: the 160-problem rank sweep of configs[3] (latency:
about one problem per SM) and the 1280-problem batch of one trial (throughput), and prints a checksum of the results.

    python tools/time_nmf.py -- lib1.so lib2.so ...
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    from bench import nmf_envelopes
    from muscle_synergies_b200 import _native as nat
    from muscle_synergies_b200 import analysis

    libs = sys.argv[sys.argv.index("--") + 1:]
    X = nmf_envelopes()
    for path in libs:
        nat.LIB_PATH = os.path.abspath(path)  # tools only: the product loads the library next to the package
        nat._lib = None
        for label, cycles, iters in (("160 problems x 2000 it", 1, 2000), ("1280 problems x 200 it", 8, 200)):
            Xs = np.stack([X * (1.0 + 0.01 * c) for c in range(cycles)])
            ranks = [k for _ in range(cycles) for k in range(1, 9) for _ in range(20)]
            seeds = [r for _ in range(cycles) for _ in range(1, 9) for r in range(20)]
            xi = [c for c in range(cycles) for _ in range(160)]
            Xd = torch.from_numpy(Xs).cuda()

            lib = nat.lib()
            real = lib.ms_nmf_mu_batched_planned
            times = []

            def timed_entry(*a):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = real(*a)
                e1.record()
                times.append((e0, e1))
                return rc

            lib.ms_nmf_mu_batched_planned = timed_entry
            for _ in range(2):
                analysis.nmf_mu_batched(Xd, ranks, seeds, max_iter=20, tol=0.0, x_index=xi)
            times.clear()
            for _ in range(5):
                res = analysis.nmf_mu_batched(Xd, ranks, seeds, max_iter=iters, tol=0.0, x_index=xi)
            torch.cuda.synchronize()
            lib.ms_nmf_mu_batched_planned = real
            kernel = sorted(a.elapsed_time(b) for a, b in times)[2] * 1e-3
            print(f"{os.path.basename(path):32s} {label:24s} {len(ranks) * iters / kernel / 1e6:8.1f} M it/s  kernel {kernel * 1e3:7.3f} ms  "
                  f"err sum {float(res.err.sum()):.6f} vaf sum {float(res.vaf[:, 0].sum()):.6f}")


if __name__ == "__main__":
    main()
