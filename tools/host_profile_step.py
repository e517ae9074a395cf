"""cProfile of the host side of a bench step (load_device + Segmenter + phase_cuts on an HBM-resident trial)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import muscle_synergies_b200 as ms
from muscle_synergies_b200.segment import Segmenter
from tools.synth_vicon import synth_layout

layout = sys.argv[1] if len(sys.argv) > 1 else "T10"
blob = synth_layout(layout, seed=5)
loader = ms.ViconLoader()
n = blob.nbytes
d = torch.empty(loader.padded_size(n), dtype=torch.uint8, device="cuda")
d[:n].copy_(torch.from_numpy(blob))


def step():
    data = loader.load_device(d, n=n, name=layout, defer_check=True)
    seg = Segmenter(data, cut_phases_of=(data.emg,))
    return seg.phase_cuts(data.emg)


for _ in range(10):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
