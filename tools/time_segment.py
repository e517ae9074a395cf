"""Times ms_segment_trial (bitmaps + transition search + window plans, one launch) for several builds of the library.

    python tools/time_segment.py LAYOUT -- lib1.so lib2.so ...
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import muscle_synergies_b200 as ms
    from muscle_synergies_b200 import _native as nat
    from muscle_synergies_b200.segment import _fz_tensor
    from tools.synth_vicon import synth_layout

    sep = sys.argv.index("--")
    layout = sys.argv[1] if sep > 1 else "T10"
    libs = sys.argv[sep + 1:]
    data = ms.load_vicon_bytes(synth_layout(layout, seed=5), name=layout)
    left, right = (_fz_tensor(fp).contiguous() for fp in data.forcepl)
    n = int(left.numel())
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    for path in libs:
        L = ctypes.CDLL(os.path.abspath(path))
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
        L.ms_transitions_workspace_bytes.restype = i64
        L.ms_transitions_workspace_bytes.argtypes = [i64]
        L.ms_segment_trial.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, i32, vp]
        work = torch.empty(int(L.ms_transitions_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
        trans = torch.zeros(40, dtype=torch.int64, device="cuda")
        loaded = torch.zeros(40, dtype=torch.int32, device="cuda")
        found = torch.zeros(1, dtype=torch.int32, device="cuda")

        def call():
            rc = L.ms_segment_trial(left.data_ptr(), right.data_ptr(), n, 10, 40, work.data_ptr(), trans.data_ptr(),
                                    loaded.data_ptr(), found.data_ptr(), None, 0, sptr)
            assert rc == 0, rc

        for _ in range(3):
            call()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in evs:
            a.record(stream)
            call()
            b.record(stream)
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        print(f"{os.path.basename(path):40s} n {n}  median {ts[10] * 1e3:.1f} us  best {ts[0] * 1e3:.1f} us  found {int(found.item())} "
              f"first {trans[:4].tolist()} sum {int(trans.sum().item())}")


if __name__ == "__main__":
    main()
