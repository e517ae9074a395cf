"""Times the single-pass loader kernel against the two-pass path on a device-resident trial (CUDA events on the
launching stream) and checks that both produce the same bits.

    python tools/time_fused.py [layout] [tile_bytes ...]
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    import muscle_synergies_b200 as ms
    from muscle_synergies_b200 import _native as nat
    from muscle_synergies_b200.vicon_data import loader as loader_mod
    from tools.synth_vicon import synth_layout

    layout = sys.argv[1] if len(sys.argv) > 1 else "T10"
    tiles = [int(a) for a in sys.argv[2:]] or [0]
    blob = synth_layout(layout, seed=1000)
    n = int(blob.nbytes)
    dev = torch.device("cuda:0")
    loader = ms.ViconLoader(dev)
    d_bytes = torch.zeros(loader.padded_size(n), dtype=torch.uint8, device=dev)
    d_bytes[:n] = torch.from_numpy(blob).to(dev)
    lib = nat.lib()
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)

    loader_mod.FORCE_PATH = "two_pass"
    ref = loader.load_device(d_bytes, n=n, name=layout)
    ref_blocks = [b.tensor[:, : b.n_rows].clone() for b in ref.blocks]
    n_kept = sum(int(b.numel()) for b in ref_blocks)
    b_alg = n + 8 * n_kept
    print(json.dumps({"layout": layout, "csv_bytes": n, "kept_doubles": n_kept, "algorithmic_bytes": b_alg}))

    def timed(call, reps=20):
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record(stream)
            call()
            b.record(stream)
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        return ts[len(ts) // 2], ts[0]

    rows = [b.n_rows for b in ref.blocks]
    keep = [int(b.tensor.shape[0]) for b in ref.blocks]
    cap1, cap2 = rows[0] + 64, rows[1] + 64
    arena = torch.empty(keep[0] * cap1 + 2 + keep[1] * cap2, dtype=torch.float64, device=dev)
    d_res = torch.empty(256 + 2 * nat.MS_LOAD_PEEK, dtype=torch.uint8, device=dev)
    for tile in tiles:
        ws_bytes = int(lib.ms_load_workspace_bytes(n, tile))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        plan = nat.LoadPlan(arena.data_ptr(), arena.numel(), (ctypes.c_int64 * 2)(cap1, cap2), tile, 0)

        def call():
            nat.check(lib.ms_load_fused(d_bytes.data_ptr(), n, ctypes.byref(plan), ws.data_ptr(), ws_bytes, d_res.data_ptr(),
                                        d_res.data_ptr() + 256, sptr), "ms_load_fused")

        arena.fill_(-7.0)
        med, best = timed(call)
        res = nat.LoadResult.from_buffer_copy(d_res[:152].cpu().numpy().tobytes())
        same = True
        for s in (0, 1):
            k, st, off = int(res.n_keep[s]), int(res.stride[s]), int(res.out_offset[s])
            got = arena[off : off + k * st].view(k, st)[: keep[s], : rows[s]]
            same = same and bool((got.view(torch.int64) == ref_blocks[s].view(torch.int64)).all())
        print(json.dumps({"kernel": "ms_load_kernel (+3 memsets)", "tile_bytes": tile, "ms_median": med, "ms_best": best,
                          "csv_gbs": n / med / 1e6, "algorithmic_gbs": b_alg / med / 1e6, "flags": res.flags,
                          "status_ok": res.status == nat.MS_ERR_NONE, "data_rows": list(res.data_rows),
                          "bits_equal_two_pass": same}))

    # the two-pass kernels on the same buffer
    src = loader_mod._Source(d_bytes, n, None)
    summary, ws2 = loader._scan(src)
    plan2 = loader_mod._plan(src, summary, layout)
    sections = (nat.Section * nat.MS_MAX_SECTIONS)()
    blocks = []
    k = 0
    for lay, (r0, r1) in zip(plan2.layouts, plan2.data_rows):
        blk = torch.empty((lay.n_keep, r1 - r0), dtype=torch.float64, device=dev)
        blocks.append(blk)
        s = sections[k]
        s.row_begin, s.row_end, s.num_cols, s.n_keep, s.d_out, s.stride = r0, r1, lay.num_cols, lay.n_keep, blk.data_ptr(), r1 - r0
        k += 1
    d_status = torch.empty(1, dtype=torch.int64, device=dev)
    d_summary = torch.empty(ctypes.sizeof(nat.ScanSummary), dtype=torch.uint8, device=dev)
    t_parse = timed(lambda: lib.ms_parse(d_bytes.data_ptr(), n, ws2.data_ptr(), sections, k, d_status.data_ptr(), sptr))
    t_scan = timed(lambda: lib.ms_scan(d_bytes.data_ptr(), n, ws2.data_ptr(), ws2.numel(), d_summary.data_ptr(), sptr))
    print(json.dumps({"two_pass": {"ms_parse": t_parse[0], "ms_scan+resolve": t_scan[0],
                                   "algorithmic_gbs_both": b_alg / (t_parse[0] + t_scan[0]) / 1e6}}))


if __name__ == "__main__":
    main()
