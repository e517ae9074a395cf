"""A/B timing of ms_scan / ms_parse for several builds of the library (tools only).
usage: python tools/ab_parse.py LAYOUT lib1.so lib2.so ..."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from muscle_synergies_b200 import _native
from tools.synth_vicon import synth_layout

layout = sys.argv[1]
blob = synth_layout(layout, seed=5)
n = blob.nbytes
d = torch.empty((n + 15) // 16 * 16 + 16, dtype=torch.uint8, device="cuda")
d[:n].copy_(torch.from_numpy(blob))
import muscle_synergies_b200 as ms
from muscle_synergies_b200.vicon_data import loader as lm
loader = ms.ViconLoader()
src = lm._Source(d, n, None)
summary, ws0 = loader._scan(src)
plan = lm._plan(src, summary, layout)
stream = torch.cuda.current_stream()
sptr = ctypes.c_void_p(stream.cuda_stream)
for path in sys.argv[2:]:
    L = ctypes.CDLL(os.path.abspath(path))
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.ms_workspace_bytes.restype = i64; L.ms_workspace_bytes.argtypes = [i64]
    L.ms_scan.argtypes = [vp, i64, vp, i64, vp, vp]
    L.ms_parse.argtypes = [vp, i64, vp, ctypes.POINTER(_native.Section), i32, vp, vp]
    ws = torch.empty(int(L.ms_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
    dsum = torch.empty(256, dtype=torch.uint8, device="cuda")
    secs = (_native.Section * 4)()
    blocks = []
    k = 0
    for lay, (r0, r1) in zip(plan.layouts, plan.data_rows):
        blk = torch.empty((lay.n_keep, r1 - r0), dtype=torch.float64, device="cuda")
        blocks.append(blk)
        s = secs[k]
        s.row_begin, s.row_end, s.num_cols, s.n_keep, s.d_out, s.stride = r0, r1, lay.num_cols, lay.n_keep, blk.data_ptr(), r1 - r0
        k += 1
    dst = torch.zeros(2, dtype=torch.int64, device="cuda")
    def t(call, reps=20):
        for _ in range(3): call()
        torch.cuda.synchronize()
        e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in e:
            a.record(stream); call(); b.record(stream)
        torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in e)[reps // 2]
    ts = t(lambda: L.ms_scan(d.data_ptr(), n, ws.data_ptr(), ws.numel(), dsum.data_ptr(), sptr))
    tp = t(lambda: L.ms_parse(d.data_ptr(), n, ws.data_ptr(), secs, k, dst.data_ptr(), sptr))
    chk = [int(b.view(torch.int64).sum().item()) for b in blocks]
    dst[1] = 0
    L.ms_parse(d.data_ptr(), n, ws.data_ptr(), secs, k, dst.data_ptr(), sptr)
    torch.cuda.synchronize()
    print(f"{os.path.basename(path):28s} scan+resolve {ts:.3f} ms  parse {tp:.3f} ms  status {int(dst[0].item())} aux {int(dst[1].item())} checksum {chk}")
