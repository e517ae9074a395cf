"""Times the drop-in call load_vicon_file(path) on a T10 file in tmpfs (page cache warm)."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import muscle_synergies_b200 as ms
from tools.synth_vicon import synth_layout
blob = synth_layout(sys.argv[1] if len(sys.argv) > 1 else "T10", seed=5)
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
p = os.path.join(d, "t.csv"); blob.tofile(p)
try:
    for _ in range(2):
        ms.load_vicon_file(p)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        data = ms.load_vicon_file(p)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(f"load_vicon_file: {dt * 1e3:.1f} ms per {blob.nbytes / 1e6:.0f} MB file = {blob.nbytes / dt / 1e9:.1f} GB/s (arrays in HBM)")
    t = time.perf_counter()
    data = ms.load_vicon_file(p); _ = data.emg.df; _ = data.traj[0].df
    print(f"  + DataFrames of both sections on the host: {(time.perf_counter() - t) * 1e3:.1f} ms")
finally:
    shutil.rmtree(d, ignore_errors=True)
