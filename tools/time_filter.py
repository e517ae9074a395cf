"""Times the EMG envelope kernels on a T10-sized recording (16 channels x 1.2 M samples)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from muscle_synergies_b200 import emg

x = torch.randn((16, 1_200_000), dtype=torch.float64, device="cuda") * 0.01
mean = emg.channel_means(x)
sos = emg.filter_coeffs(6.0, 2000, 4)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("channel_means ms", t(lambda: emg.channel_means(x)))
print("rms_envelope ms", t(lambda: emg.rms_envelope(x, 1000, mean)))
print("sosfiltfilt(order 4) ms", t(lambda: emg.sos_filter(x, sos, True, mean, True)))
print("sosfilt(order 4) ms", t(lambda: emg.sos_filter(x, sos, False, mean, True)))
