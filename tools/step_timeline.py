"""Where the host's time goes in one step (HBM-resident CSV -> arrays + transitions + 32 EMG windows): wall-clock marks
inside the single-pass load and around the segmenter, averaged over a few steps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import muscle_synergies_b200 as ms
from muscle_synergies_b200.segment import Segmenter
from muscle_synergies_b200.vicon_data import loader as loader_mod
from tools.synth_vicon import synth_layout

layout = sys.argv[1] if len(sys.argv) > 1 else "T10"
blob = synth_layout(layout, seed=5)
loader = ms.ViconLoader()
n = blob.nbytes
d = torch.empty(loader.padded_size(n), dtype=torch.uint8, device="cuda")
d[:n].copy_(torch.from_numpy(blob))


def step(marks=None):
    loader_mod.TIMELINE = marks
    if marks is not None:
        marks.append(("step", time.perf_counter()))
    data = loader.load_device(d, n=n, name=layout, defer_check=True)
    if marks is not None:
        marks.append(("loaded", time.perf_counter()))
    seg = Segmenter(data, cut_phases_of=(data.emg,))
    if marks is not None:
        marks.append(("segmented", time.perf_counter()))
    cuts = seg.phase_cuts(data.emg)
    if marks is not None:
        marks.append(("end", time.perf_counter()))
    return cuts


for _ in range(5):
    step()
torch.cuda.synchronize()
acc = {}
reps = 30
for _ in range(reps):
    marks = []
    step(marks)
    torch.cuda.synchronize()
    for (a, ta), (b, tb) in zip(marks, marks[1:]):
        acc[f"{a} -> {b}"] = acc.get(f"{a} -> {b}", 0.0) + (tb - ta)
total = 0.0
for k, v in acc.items():
    print(f"{k:32s} {v / reps * 1e6:8.1f} us")
    total += v
print(f"{'total':32s} {total / reps * 1e6:8.1f} us")
