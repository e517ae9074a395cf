"""Small driver for ncu: loads one synthetic trial a few times (scan + parse + segment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import muscle_synergies_b200 as ms
from muscle_synergies_b200.segment import Segmenter
from tools.synth_vicon import synth_layout

layout = sys.argv[1] if len(sys.argv) > 1 else "T127"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
blob = synth_layout(layout, seed=5)
loader = ms.ViconLoader()
n = blob.nbytes
d = torch.empty(loader.padded_size(n), dtype=torch.uint8, device="cuda")
d[:n].copy_(torch.from_numpy(blob))
for _ in range(reps):
    data = loader.load_device(d, n=n, name=layout, defer_check=True)
    seg = Segmenter(data, cut_phases_of=(data.emg,))
    cuts = seg.phase_cuts(data.emg)
torch.cuda.synchronize()
print("done", layout, n)
