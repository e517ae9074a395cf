"""EMG envelope chain on the GPU (SURVEY.md section 8f, rank 1 "next" row).

Mirrors of the reference's preprocessing functions (src/muscle_synergies/analysis.py):
`zero_center` (:230-249), `linear_envelope` (:252-311), `digital_filter` (:314-432), `rms`
(:435-507), `normalize` (:510-525), `time_normalize` (:551-594, kind="linear"), same signatures
on DataFrames, plus `envelope_windows`, which keeps
everything in HBM between the cut windows and the NMF stage: trial-wide zero-centred moving RMS,
then per window linear time-normalisation to `reduce_to` samples and division by the column
maximum - the flow of docs/source/tutorials "Finding muscle synergies" (cells 10-23) applied
per gait cycle.  float64; parity with the reference functions is to a tolerance (sums are
ordered differently), see tests/test_emg_gpu.py.
"""
import ctypes
from typing import Optional, Sequence, Union

import numpy as np
import pandas

from . import _native as nat


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise nat.NativeError("muscle_synergies_b200.emg needs a CUDA device; there is no CPU fallback")
    return torch


def _stream(torch, dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _channel_major(signal_df: pandas.DataFrame):
    torch = _torch()
    arr = np.array(signal_df.to_numpy(dtype=np.float64).T, order="C", copy=True)  # (channels, rows)
    return torch.from_numpy(arr).cuda()


@nat.on_tensor_device
def channel_means(x):
    """Mean of every row of a (channels, samples) float64 CUDA tensor."""
    torch = _torch()
    x = x if x.stride(1) == 1 else x.contiguous()
    out = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
    nat.check(nat.lib().ms_channel_means(x.data_ptr(), int(x.stride(0)), int(x.shape[0]), int(x.shape[1]), out.data_ptr(),
                                         _stream(torch, x.device)), "ms_channel_means")
    return out


@nat.on_tensor_device
def rms_envelope(x, window: int, mean=None):
    """sqrt(convolve((x - mean)^2, ones(window)/window, "same")) per row of a (channels, samples) tensor."""
    torch = _torch()
    x = x if x.stride(1) == 1 else x.contiguous()
    n_ch, n = int(x.shape[0]), int(x.shape[1])
    if window < 1 or window > n:
        raise ValueError("window must be between 1 and the number of samples")
    out = torch.empty((n_ch, n), dtype=torch.float64, device=x.device)
    nat.check(
        nat.lib().ms_rms_envelope(x.data_ptr(), int(x.stride(0)), n_ch, n, mean.data_ptr() if mean is not None else None,
                                  int(window), out.data_ptr(), n, _stream(torch, x.device)),
        "ms_rms_envelope",
    )
    return out


def filter_coeffs(critical_freqs, sampling_frequency: int, order: int, filter_type: str = "butter",
                  band_type: str = "lowpass", cheby_param: Optional[float] = None) -> np.ndarray:
    """Second-order sections of the filter, designed on the host by the scipy calls the reference
    makes (analysis.py:386-409): a handful of numbers; the recursion over the samples is the GPU's."""
    from scipy import signal

    if filter_type not in {"butter", "cheby1", "cheby2"}:
        raise ValueError("filter type not understood.")
    if filter_type == "butter":
        return signal.butter(order, critical_freqs, btype=band_type, output="sos", fs=sampling_frequency)
    design = signal.cheby1 if filter_type == "cheby1" else signal.cheby2
    return design(order, cheby_param, critical_freqs, btype=band_type, output="sos", fs=sampling_frequency)


@nat.on_tensor_device
def sos_filter(x, sos: np.ndarray, zero_lag: bool = True, mean=None, rectify: bool = False):
    """scipy.signal.sosfiltfilt (zero_lag) or sosfilt along every row of a (channels, samples) float64
    CUDA tensor; with `mean` / `rectify` the input is |x - mean| (the linear envelope's rectifier, fused)."""
    torch = _torch()
    from scipy import signal

    x = x if x.stride(1) == 1 else x.contiguous()
    n_ch, n = int(x.shape[0]), int(x.shape[1])
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise ValueError("sos array must be 2D with shape (n_sections, 6)")
    n_sections = sos.shape[0]
    if n_sections > 8:
        raise NotImplementedError("the CUDA filter holds up to 8 second-order sections (order 16)")
    if not np.all(sos[:, 3] == 1.0):
        raise ValueError("sos[:, 3] should be all ones")
    padlen, zi = 0, None
    if zero_lag:
        # sosfiltfilt's default padding (scipy/signal/_signaltools.py): 3 * ntaps, odd extension
        ntaps = 2 * n_sections + 1
        ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
        padlen = 3 * ntaps
        if n <= padlen:
            raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
        zi = np.ascontiguousarray(signal.sosfilt_zi(sos), dtype=np.float64)
    lib = nat.lib()
    out = torch.empty((n_ch, n), dtype=torch.float64, device=x.device)
    if n_ch == 0 or n == 0:
        return out
    work = torch.empty(int(lib.ms_sosfilt_workspace_bytes(n, n_ch, padlen, int(zero_lag))), dtype=torch.uint8, device=x.device)
    dptr = ctypes.POINTER(ctypes.c_double)
    nat.check(
        lib.ms_sosfilt(x.data_ptr(), int(x.stride(0)), n_ch, n, sos.ctypes.data_as(dptr), n_sections,
                       zi.ctypes.data_as(dptr) if zi is not None else None, padlen, int(zero_lag),
                       mean.data_ptr() if mean is not None else None, int(bool(rectify)), out.data_ptr(), n,
                       work.data_ptr(), _stream(torch, x.device)),
        "ms_sosfilt",
    )
    return out


@nat.on_tensor_device
def time_normalize_windows(env, starts: Sequence[int], stops: Sequence[int], reduce_to: int, normalize: bool = True):
    """(n_windows, reduce_to, channels) float64 CUDA tensor from a (channels, samples) envelope."""
    torch = _torch()
    env = env if env.stride(1) == 1 else env.contiguous()
    n_ch = int(env.shape[0])
    # through pinned memory, without waiting: a copy from pageable memory would wait for everything queued on the stream
    meta = torch.tensor([list(starts), list(stops)], dtype=torch.int64, pin_memory=True).to(env.device, non_blocking=True)
    out = torch.empty((len(starts), reduce_to, n_ch), dtype=torch.float64, device=env.device)
    nat.check(
        nat.lib().ms_time_normalize_windows(env.data_ptr(), int(env.stride(0)), n_ch, meta[0].data_ptr(), meta[1].data_ptr(),
                                            len(starts), int(reduce_to), int(bool(normalize)), out.data_ptr(),
                                            _stream(torch, env.device)),
        "ms_time_normalize_windows",
    )
    return out


def envelope_windows(device, windows, window_size: float = 0.5, reduce_to: int = 200, normalize: bool = True,
                     method: str = "rms", critical_freqs=None, order: int = 4, filter_type: str = "butter",
                     cheby_param: Optional[float] = None):
    """Trial-wide envelope of a DeviceData (EMG), then every (frame, subframe) window resampled to
    `reduce_to` points and amplitude-normalised.  Stays on the GPU.

    method="rms": zero-centred moving RMS over `window_size` seconds (the tutorial's choice);
    method="linear_envelope": zero-centre, rectify, zero-lag low-pass at `critical_freqs` Hz.
    Returns (n_windows, reduce_to, n_muscles) float64 - each [w] is the X of one NMF problem."""
    x = device.tensor
    n = int(x.shape[1])
    mean = channel_means(x)
    if method == "rms":
        env = rms_envelope(x, round(window_size * device.sampling_frequency), mean)
    elif method == "linear_envelope":
        if critical_freqs is None:
            raise ValueError("linear_envelope needs critical_freqs (the low-pass cut-off in Hz)")
        sos = filter_coeffs(critical_freqs, device.sampling_frequency, order, filter_type, "lowpass", cheby_param)
        env = sos_filter(x, sos, zero_lag=True, mean=mean, rectify=True)
    else:
        raise ValueError('method must be "rms" or "linear_envelope"')
    idx = [device.to_index(w) for w in windows]
    starts = [s.indices(n)[0] for s in idx]
    stops = [max(s.indices(n)[0], s.indices(n)[1]) for s in idx]
    return time_normalize_windows(env, starts, stops, reduce_to, normalize)


# ---- DataFrame-level mirrors of the reference API ----------------------------------------------------------
def _like(signal_df: pandas.DataFrame, inplace: bool, values: np.ndarray) -> pandas.DataFrame:
    if inplace:
        signal_df.iloc[:, :] = values
        return signal_df
    return pandas.DataFrame(values, index=signal_df.index, columns=signal_df.columns)


def zero_center(signal_df: pandas.DataFrame, inplace: bool = False) -> pandas.DataFrame:
    x = _channel_major(signal_df)
    centred = x - channel_means(x)[:, None]
    return _like(signal_df, inplace, centred.T.cpu().numpy())


def digital_filter(signal_df: pandas.DataFrame, critical_freqs, sampling_frequency: int, order: int,
                   filter_type: str = "butter", band_type: str = "lowpass", zero_lag: bool = True,
                   cheby_param: Optional[float] = None, inplace: bool = False) -> pandas.DataFrame:
    sos = filter_coeffs(critical_freqs, sampling_frequency, order, filter_type, band_type, cheby_param)
    out = sos_filter(_channel_major(signal_df), sos, zero_lag=zero_lag)
    return _like(signal_df, inplace, out.T.cpu().numpy())


def linear_envelope(signal_df: pandas.DataFrame, critical_freqs, sampling_frequency: int, order: int,
                    filter_type: str = "butter", zero_lag: bool = True, cheby_param: Optional[float] = None,
                    zero_center_: bool = True, inplace: bool = False) -> pandas.DataFrame:
    sos = filter_coeffs(critical_freqs, sampling_frequency, order, filter_type, "lowpass", cheby_param)
    x = _channel_major(signal_df)
    out = sos_filter(x, sos, zero_lag=zero_lag, mean=channel_means(x) if zero_center_ else None, rectify=True)
    return _like(signal_df, inplace, out.T.cpu().numpy())


def rms(signal_df: pandas.DataFrame, window_size: Union[int, float], inplace: bool = False,
        sampling_frequency: Optional[int] = None) -> pandas.DataFrame:
    if sampling_frequency is not None:
        window_size = round(window_size * sampling_frequency)
    out = rms_envelope(_channel_major(signal_df), int(window_size))
    return _like(signal_df, inplace, out.T.cpu().numpy())


def normalize(signal_df: pandas.DataFrame, inplace: bool = False) -> pandas.DataFrame:
    x = _channel_major(signal_df)
    out = x / x.abs().amax(dim=1, keepdim=True)
    return _like(signal_df, inplace, out.T.cpu().numpy())


def time_normalize(signal_df: pandas.DataFrame, reduce_to: int, kind="linear", fill_value="extrapolate") -> pandas.DataFrame:
    if kind != "linear" or fill_value != "extrapolate":
        # what the reference does for every kind: scipy.interpolate.interp1d (analysis.py:584-594); the CUDA kernel
        # covers the linear case the pipeline uses
        from scipy import interpolate

        percent_domain = np.linspace(0, 1, signal_df.shape[0])
        interp_func = interpolate.interp1d(percent_domain, signal_df, axis=0, copy=False, kind=kind, fill_value=fill_value)
        desired_domain = np.linspace(0, 1, reduce_to)
        return pandas.DataFrame(interp_func(desired_domain), index=desired_domain, columns=signal_df.columns)
    x = _channel_major(signal_df)
    out = time_normalize_windows(x, [0], [int(x.shape[1])], reduce_to, normalize=False)[0]
    return pandas.DataFrame(out.cpu().numpy(), index=np.linspace(0, 1, reduce_to), columns=signal_df.columns)
