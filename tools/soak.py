"""Soak: many steps back to back; device memory and host RSS must stay flat (tools only)."""
import os, sys, time, resource
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import muscle_synergies_b200 as ms
from muscle_synergies_b200.segment import Segmenter
from muscle_synergies_b200.pipeline import trial_synergies
from tools.synth_vicon import synth_layout

blob = synth_layout("T127", seed=5)
n = blob.nbytes
loader = ms.ViconLoader()
d = torch.empty(loader.padded_size(n), dtype=torch.uint8, device="cuda")
d[:n].copy_(torch.from_numpy(blob))
pinned = torch.empty(loader.padded_size(n), dtype=torch.uint8, pin_memory=True)
pinned.numpy()[:n] = blob


def snap(tag):
    torch.cuda.synchronize()
    print(f"{tag}: device reserved {torch.cuda.memory_reserved() / 1e6:.0f} MB, allocated {torch.cuda.memory_allocated() / 1e6:.0f} MB, "
          f"host RSS {resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e3:.0f} MB")


for rnd in range(3):
    t = time.perf_counter()
    for _ in range(300):
        data = loader.load_device(d, n=n, name="soak", defer_check=True)
        seg = Segmenter(data, cut_phases_of=(data.emg,))
        cuts = seg.phase_cuts(data.emg)
    for _ in range(20):
        trial_synergies(loader.load_device(d, n=n, name="soak"), 1, 4, n_restarts=4, max_iter=50)
    k = 0
    for data in loader.load_many([pinned[:n]] * 40, to_host=True):
        k += 1
    del data, seg, cuts
    snap(f"round {rnd} ({time.perf_counter() - t:.1f} s)")
