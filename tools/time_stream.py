"""Step time of a stream of HBM-resident trials through load_device_many vs one trial at a time (tools only).

    python tools/time_stream.py [LAYOUT] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import muscle_synergies_b200 as ms
from muscle_synergies_b200.segment import Segmenter
from tools.synth_vicon import synth_layout

layout = sys.argv[1] if len(sys.argv) > 1 else "T10"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
blob = synth_layout(layout, seed=5)
loader = ms.ViconLoader()
n = blob.nbytes
d = torch.empty(loader.padded_size(n), dtype=torch.uint8, device="cuda")
d[:n].copy_(torch.from_numpy(blob))


def consume(data):
    seg = Segmenter(data, cut_phases_of=(data.emg,))
    return seg, seg.phase_cuts(data.emg)


def serial(k):
    for _ in range(k):
        consume(loader.load_device(d, n=n, name=layout, defer_check=True))


def streamed(k, depth, hi=None):
    for data in loader.load_device_many(((d, n) for _ in range(k)), depth=depth, stream=hi):
        if hi is None:
            consume(data)
        else:
            with torch.cuda.stream(hi):
                consume(data)
    if hi is not None:
        torch.cuda.current_stream().wait_stream(hi)


def alloc_counts():
    st = torch.cuda.memory_stats()
    return st.get("num_device_alloc", 0), st.get("num_device_free", 0), st.get("num_alloc_retries", 0), torch.cuda.memory_reserved() >> 20


def per_step(k, depth, hi):
    """host wall time between yields: where a slow step is"""
    import time
    ts = [time.perf_counter()]
    for data in loader.load_device_many(((d, n) for _ in range(k)), depth=depth, stream=hi):
        with torch.cuda.stream(hi):
            consume(data)
        ts.append(time.perf_counter())
    dt = sorted((b - a) * 1e3 for a, b in zip(ts, ts[1:]))
    return dt[len(dt) // 2], dt[-1], dt[-3:]


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


ref = loader.load_device(d, n=n, name=layout)
ref_seg, ref_cuts = consume(ref)
serial(5)
print(f"serial               {timed(lambda: serial(steps)):.4f} ms/step")
hi = loader.work_stream
for depth in (1, 2, 3):
    streamed(5, depth)
    print(f"streamed depth {depth}     {timed(lambda: streamed(steps, depth)):.4f} ms/step")
    streamed(5, depth, hi)
    print(f"streamed depth {depth} hi  {timed(lambda: streamed(steps, depth, hi)):.4f} ms/step")
for rep in range(4):
    a0 = alloc_counts()
    t = timed(lambda: streamed(steps, 2, hi))
    print(f"rep {rep}: streamed depth 2 hi {t:.4f} ms/step   device allocs/frees/retries/reserved MB before {a0} after {alloc_counts()}")
print("per-step wall (median, max, top3):", per_step(steps, 2, hi))
# same results
for data in loader.load_device_many(((d, n) for _ in range(3)), depth=2):
    seg, cuts = consume(data)
    assert list(seg.transitions) == list(ref_seg.transitions)
    for a, b in zip(data.blocks, ref.blocks):
        assert a.n_rows == b.n_rows
        assert torch.equal(a.tensor[:, : a.n_rows].contiguous().view(torch.int64), b.tensor[:, : b.n_rows].contiguous().view(torch.int64))
    for a, b in zip(cuts, ref_cuts):
        ta = a.tensor if hasattr(a, "tensor") else a
        tb = b.tensor if hasattr(b, "tensor") else b
        if isinstance(ta, torch.Tensor):
            assert torch.equal(ta.view(torch.int64), tb.view(torch.int64))
print("streamed results identical to load_device:", loader.stats)
