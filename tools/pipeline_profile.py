"""Where the per-trial time of trial_synergies goes (tools only)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import muscle_synergies_b200 as ms
from muscle_synergies_b200.pipeline import trial_synergies
from tools.synth_vicon import synth_layout

layout = sys.argv[1] if len(sys.argv) > 1 else "T10"
blob = synth_layout(layout, seed=5)
n = blob.nbytes
loader = ms.ViconLoader()
d = torch.empty(loader.padded_size(n), dtype=torch.uint8, device="cuda")
d[:n].copy_(torch.from_numpy(blob))


def one():
    data = loader.load_device(d, n=n, name=layout)
    return trial_synergies(data, 1, 8, n_restarts=20, random_state=0, max_iter=200, tol=0.0)


for _ in range(3):
    one()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    one()
torch.cuda.synchronize()
print("ms/trial", (time.perf_counter() - t) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    one()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
