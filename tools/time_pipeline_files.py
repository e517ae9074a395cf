"""configs[4] on one GPU: synergies_for_files over distinct trial files in tmpfs; wall clock per trial, with and
without the one-deep overlap (tools only)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import muscle_synergies_b200 as ms
from muscle_synergies_b200 import pipeline
from tools.synth_vicon import synth_layout

layout = sys.argv[1] if len(sys.argv) > 1 else "T127"
n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 8
root = "/dev/shm/ms_b200_tpf"
os.makedirs(root, exist_ok=True)
paths = []
for i in range(n_files):
    p = os.path.join(root, f"t{i}.csv")
    synth_layout(layout, seed=3000 + i).tofile(p)
    paths.append(p)
loader = ms.ViconLoader()
kw = dict(min_components=1, max_components=8, n_restarts=20, random_state=0, max_iter=200, tol=0.0)


def run(serial):
    t = time.perf_counter()
    if serial:
        for path, data in loader.load_files(paths, to_host=False):
            pipeline.trial_synergies(data, **kw)
    else:
        for _ in pipeline.synergies_for_files(paths, loader=loader, **kw):
            pass
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n_files * 1e3


run(False)
run(True)
for _ in range(2):
    print("overlapped ms/trial", round(run(False), 2), " serial ms/trial", round(run(True), 2))
t = time.perf_counter()
for path, data in loader.load_files(paths, to_host=False):
    pass
torch.cuda.synchronize()
print("load_files only ms/trial", round((time.perf_counter() - t) / n_files * 1e3, 2))
pr = cProfile.Profile()
pr.enable()
run(False)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
import shutil

shutil.rmtree(root, ignore_errors=True)
