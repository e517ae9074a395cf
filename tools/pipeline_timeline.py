"""Host wall time of each stage of synergies_for_files' loop (tools only): where the host waits for the GPU."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import muscle_synergies_b200 as ms
from muscle_synergies_b200 import pipeline
from muscle_synergies_b200.segment import Segmenter
from tools.synth_vicon import synth_layout

layout = sys.argv[1] if len(sys.argv) > 1 else "T127"
n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 8
root = "/dev/shm/ms_b200_tpl"
os.makedirs(root, exist_ok=True)
paths = []
for i in range(n_files):
    p = os.path.join(root, f"t{i}.csv")
    synth_layout(layout, seed=3000 + i).tofile(p)
    paths.append(p)
loader = ms.ViconLoader()
kw = dict(min_components=1, max_components=8, n_restarts=20, random_state=0, max_iter=200, tol=0.0)
work = loader.work_stream


def run(report):
    acc = {"next(load_files)": 0.0, "Segmenter": 0.0, "trial_synergies(defer)": 0.0, "finish(prev)": 0.0}
    t_all = time.perf_counter()
    it = iter(loader.load_files(paths, to_host=False))
    prev = None
    while True:
        t = time.perf_counter()
        try:
            path, data = next(it)
        except StopIteration:
            break
        acc["next(load_files)"] += time.perf_counter() - t
        t = time.perf_counter()
        with torch.cuda.stream(work):
            for blk in data.blocks:
                blk.tensor.record_stream(work)
            seg = Segmenter(data)
        acc["Segmenter"] += time.perf_counter() - t
        t = time.perf_counter()
        cur = pipeline.trial_synergies(data, defer=True, segmenter=seg, **kw)
        acc["trial_synergies(defer)"] += time.perf_counter() - t
        t = time.perf_counter()
        if prev is not None:
            prev.finish()
        acc["finish(prev)"] += time.perf_counter() - t
        prev = cur
    prev.finish()
    torch.cuda.synchronize()
    total = time.perf_counter() - t_all
    if report:
        for k, v in acc.items():
            print(f"{k:26s} {v / n_files * 1e3:7.2f} ms/trial")
        print(f"{'total':26s} {total / n_files * 1e3:7.2f} ms/trial")


run(False)
run(False)
for _ in range(4):
    run(True)
# GPU time of one trial's NMF + envelopes alone
data = loader.load_file(paths[0])
seg = Segmenter(data)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    pipeline.trial_synergies(data, defer=True, segmenter=seg, **kw).finish()
e1.record()
torch.cuda.synchronize()
print(f"GPU time envelopes + NMF per trial (one after the other, finish() included): {e0.elapsed_time(e1) / 5:.2f} ms")
import shutil

shutil.rmtree(root, ignore_errors=True)
