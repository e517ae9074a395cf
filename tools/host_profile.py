"""cProfile of the host side of one step (HBM-resident CSV -> arrays + segments), to find the
Python overhead around the kernels."""
import cProfile, pstats, sys, os, time
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

for _ in range(5):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(30):
    step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t) / 30 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
