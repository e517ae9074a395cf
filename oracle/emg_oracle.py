"""CPU oracle for the EMG envelope chain ("next" row).
TEST INFRASTRUCTURE ONLY.

Each function makes the same numpy / scipy calls as the reference
(src/muscle_synergies/analysis.py): zero_center :230-249 (df - df.mean()), rms :435-507
(np.sqrt(np.convolve(sq, ones(w)/w, "same")) per column), normalize :510-525
(df / abs(df).max()), time_normalize :551-594 (scipy interp1d on linspace(0, 1, n)),
digital_filter :314-432 (scipy.signal butter/cheby1/cheby2 design as "sos", then sosfiltfilt or
sosfilt along axis 0), linear_envelope :252-311 (zero_center -> abs -> low-pass digital_filter).
The reference has no tests for these; tests/test_emg_oracle.py pins this module against the
live reference functions when /root/reference is present.
"""
import numpy as np
from scipy import interpolate


def zero_center(x):
    """x: (rows, channels).  pandas' column mean, like the reference (not numpy's pairwise sum)."""
    import pandas as pd

    df = pd.DataFrame(x)
    return (df - df.mean()).to_numpy()


def rms(x, window):
    win = (1 / float(window)) * np.ones(window)
    return np.apply_along_axis(lambda col: np.sqrt(np.convolve(col ** 2, win, "same")), 0, x)


def normalize(x):
    return x / np.abs(x).max(axis=0)


def time_normalize(x, reduce_to):
    n = x.shape[0]
    f = interpolate.interp1d(np.linspace(0, 1, n), x, axis=0, copy=False, kind="linear", fill_value="extrapolate")
    return f(np.linspace(0, 1, reduce_to))


def filter_coeffs(critical_freqs, sampling_frequency, order, filter_type="butter", band_type="lowpass", cheby_param=None):
    from scipy import signal

    if filter_type == "butter":
        return signal.butter(order, critical_freqs, btype=band_type, output="sos", fs=sampling_frequency)
    design = {"cheby1": signal.cheby1, "cheby2": signal.cheby2}[filter_type]
    return design(order, cheby_param, critical_freqs, btype=band_type, output="sos", fs=sampling_frequency)


def digital_filter(x, critical_freqs, sampling_frequency, order, filter_type="butter", band_type="lowpass",
                   zero_lag=True, cheby_param=None):
    from scipy import signal

    sos = filter_coeffs(critical_freqs, sampling_frequency, order, filter_type, band_type, cheby_param)
    return (signal.sosfiltfilt if zero_lag else signal.sosfilt)(sos, x, axis=0)


def linear_envelope(x, critical_freqs, sampling_frequency, order, filter_type="butter", zero_lag=True,
                    cheby_param=None, zero_center_=True):
    if zero_center_:
        x = zero_center(x)
    return digital_filter(np.abs(x), critical_freqs, sampling_frequency, order, filter_type, "lowpass", zero_lag,
                          cheby_param)


def envelope_windows_linear(emg, row_ranges, critical_freqs, sampling_frequency, order, reduce_to=200):
    env = linear_envelope(emg, critical_freqs, sampling_frequency, order)
    return np.stack([normalize(time_normalize(env[a:b], reduce_to)) for a, b in row_ranges])


def envelope_windows(emg, row_ranges, window, reduce_to=200):
    """emg: (rows, channels) of the whole trial; returns (n_windows, reduce_to, channels)."""
    env = rms(zero_center(emg), window)
    return np.stack([normalize(time_normalize(env[a:b], reduce_to)) for a, b in row_ranges])
