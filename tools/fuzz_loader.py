"""One-off differential fuzz: random ragged Vicon files through the CUDA loader and the Python oracle
(bit-exact arrays, same exceptions).  usage: python tools/fuzz_loader.py FIRST_SEED COUNT"""
import os, random, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
g.build()
import muscle_synergies_b200 as ms
from oracle import vicon_oracle as vo
import importlib.util
spec = importlib.util.spec_from_file_location("tl", os.path.join(ROOT, "tests", "test_loader_gpu.py"))
tl = importlib.util.module_from_spec(spec); spec.loader.exec_module(tl)

first, count = int(sys.argv[1]), int(sys.argv[2])
inject = len(sys.argv) > 3 and sys.argv[3] == "--bad"  # also plant fields float() rejects (error parity)

bad = 0
raised = 0
with tempfile.TemporaryDirectory() as d:
    for seed in range(first, first + count):
        rnd = random.Random(seed)
        blob = tl._ragged_file(rnd, rnd.randrange(1, 600), rnd.randrange(1, 80), "\n" if inject else rnd.choice(["\n", "\r\n", "\r"]))
        if inject and rnd.random() < 0.8:
            blob = tl._plant_bad_fields(blob, rnd)
        path = os.path.join(d, "f.csv")
        open(path, "wb").write(blob)
        try:
            want, werr = vo.load_vicon_file_oracle(path), None
        except Exception as e:  # noqa: BLE001
            want, werr = None, e
        try:
            got, gerr = ms.load_vicon_file(path), None
        except Exception as e:  # noqa: BLE001
            got, gerr = None, e
        if (werr is None) != (gerr is None) or (werr is not None and (type(werr), str(werr)) != (type(gerr), str(gerr))):
            bad += 1
            print("seed", seed, "exception mismatch:", repr(werr), "vs", repr(gerr))
            continue
        if werr is not None:
            raised += 1
            continue
        for dev, odev in zip(tl.all_devices(got), want.all_devices()):
            a, b = tl.bits(dev.df.to_numpy()), tl.bits(vo.device_array(odev))
            if a.shape != b.shape or not (a == b).all():
                bad += 1
                print("seed", seed, "array mismatch in", dev.name)
                break
print("fuzzed", count, "files,", raised, "of them raise in the oracle; mismatches:", bad)
