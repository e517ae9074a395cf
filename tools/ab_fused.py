"""A/B timing of ms_load_fused for several builds of the library (tools only).

    python tools/ab_fused.py --build NAME=-DFLAG[,-DFLAG2] ...     # here (nvcc): writes gpurun_variants_NAME.so
    python tools/ab_fused.py LAYOUT [tile_bytes] -- lib1.so lib2.so ...   # on the GPU box
"""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(specs):
    from muscle_synergies_b200.csrc import build as b

    for spec in specs:
        name, _, flags = spec.partition("=")
        out = os.path.join(ROOT, f"gpurun_variants_{name}.so")
        cmd = [b.NVCC] + b.FLAGS + [f for f in flags.split(",") if f] + ["-o", out] + b.sources()
        subprocess.check_call(cmd, cwd=b.HERE)
        print(out)


def main():
    if sys.argv[1] == "--build":
        return build(sys.argv[2:])
    import torch

    from muscle_synergies_b200 import _native as nat
    from tools.synth_vicon import synth_layout

    sep = sys.argv.index("--")
    layout = sys.argv[1]
    tile = int(sys.argv[2]) if sep > 2 else 0
    libs = sys.argv[sep + 1:] or sorted(glob.glob(os.path.join(ROOT, "gpurun_variants_*.so")))
    blob = synth_layout(layout, seed=1000)
    n = int(blob.nbytes)
    dev = torch.device("cuda:0")
    d = torch.zeros((n + 15) // 16 * 16 + 16, dtype=torch.uint8, device=dev)
    d[:n] = torch.from_numpy(blob).to(dev)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    arena = torch.empty(n // 8 * 2, dtype=torch.float64, device=dev)
    d_res = torch.empty(256 + 2 * nat.MS_LOAD_PEEK, dtype=torch.uint8, device=dev)
    rows_guess = n // 300
    for path in libs:
        L = ctypes.CDLL(os.path.abspath(path))
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
        L.ms_load_workspace_bytes.restype = i64
        L.ms_load_workspace_bytes.argtypes = [i64, i32]
        L.ms_load_fused.argtypes = [vp, i64, ctypes.POINTER(nat.LoadPlan), vp, i64, vp, vp, vp]
        ws_bytes = int(L.ms_load_workspace_bytes(n, tile))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        plan = nat.LoadPlan(arena.data_ptr(), arena.numel(), (ctypes.c_int64 * 2)(rows_guess, 0), tile, int(os.environ.get("MS_AB_OVERHANG", "0")))

        def call():
            rc = L.ms_load_fused(d.data_ptr(), n, ctypes.byref(plan), ws.data_ptr(), ws_bytes, d_res.data_ptr(),
                                 d_res.data_ptr() + 256, sptr)
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
        res = nat.LoadResult.from_buffer_copy(d_res[:152].cpu().numpy().tobytes())
        chk = 0
        for s in (0, 1):
            k, st, off, r = int(res.n_keep[s]), int(res.stride[s]), int(res.out_offset[s]), int(res.data_rows[s])
            chk ^= int(arena[off : off + k * st].view(k, st)[:, :r].contiguous().view(torch.int64).sum().item())
        print(f"{os.path.basename(path):40s} tile {tile:6d}  median {ts[10]:.4f} ms  best {ts[0]:.4f} ms  flags {res.flags} "
              f"status_ok {res.status == nat.MS_ERR_NONE} rows {list(res.data_rows)} checksum {chk & 0xffffffffffff:x}")


if __name__ == "__main__":
    main()
