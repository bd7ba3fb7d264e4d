"""numpy twin of scema_b200/csrc/synth.cu (include/scema_synth.h): identical bits on the host.

Every arithmetic step is a separately rounded IEEE double operation in the same order as the
device code, and the counter-based hash is pure uint64 arithmetic, so the host array equals the
device array bit-for-bit (tests/test_gpu_parity.py::test_synth_matches_numpy).
"""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)
_U = 1.1102230246251565e-16  # 2^-53


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def hash4(seed, a, b, stream):
    with np.errstate(over="ignore"):
        a = np.asarray(a, dtype=np.uint64)
        b = np.asarray(b, dtype=np.uint64)
        z = _mix64(np.uint64(seed) + _G * (a + np.uint64(1)))
        z = _mix64(z + b)
        return _mix64(z + np.uint64(stream))


def u01(x):
    return (x >> np.uint64(11)).astype(np.float64) * _U


def offsets(seed, n, cluster_size, len_min, len_max, first=0):
    q = (np.arange(n, dtype=np.uint64) + np.uint64(first)) // np.uint64(cluster_size)
    lens = np.uint64(len_min) + hash4(seed, q, 0, 7) % np.uint64(len_max - len_min + 1)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return off


def _value(seed, i, q, c, t, amp, pert):
    A = amp * (2.0 * u01(hash4(seed, q, c, 1)) - 1.0)
    beta = u01(hash4(seed, q, c, 2)) - 0.5
    delta = pert * (2.0 * u01(hash4(seed, i, c, 3)) - 1.0)
    centre = A * (t + beta * (t * t))
    return centre + delta * t


def _value_smooth(seed, i, q, c, t, amp, pert, spread):
    """model 1 of include/scema_synth.h (dogbone emulation), operation for operation as synth_value_smooth."""
    dq = spread * (2.0 * u01(hash4(seed, q, 0, 11)) - 1.0)
    zz = (amp * t) * (1.0 + dq)
    if c == 2:
        centre = zz
    elif c < 2:
        centre = -0.3 * zz
    else:
        centre = zz * (0.01 * (2.0 * u01(hash4(seed, q, c, 12)) - 1.0))
    delta = pert * (2.0 * u01(hash4(seed, i, c, 3)) - 1.0)
    return centre + delta * t


def _any(model, spread, seed, i, q, c, t, amp, pert):
    return _value_smooth(seed, i, q, c, t, amp, pert, spread) if model == 1 else _value(seed, i, q, c, t, amp, pert)


def histories(seed, n, cluster_size, amp, pert, off, first=0, model=0, spread=0.0):
    """-> steps [off[n], 6] float64 for the given (shard-relative) offsets."""
    off = np.asarray(off, dtype=np.uint64)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    total = int(off[-1])
    i = np.repeat(np.arange(n, dtype=np.uint64) + np.uint64(first), lens)
    s = np.arange(total, dtype=np.int64) - np.repeat(off[:-1].astype(np.int64), lens)
    Lm1 = np.repeat(lens - 1, lens).astype(np.float64)
    t = s.astype(np.float64) / Lm1
    q = i // np.uint64(cluster_size)
    out = np.empty((total, 6), dtype=np.float64)
    for c in range(6):
        out[:, c] = _any(model, spread, seed, i, q, c, t, amp, pert)
    return out


def rows(seed, n, cluster_size, spline_points, amp, pert, first=0, model=0, spread=0.0):
    """-> already-resampled rows [n, 6*P] in the reference's p*6+c order."""
    P = spline_points
    i = (np.arange(n, dtype=np.uint64) + np.uint64(first))[:, None]
    q = i // np.uint64(cluster_size)
    t = (np.arange(P, dtype=np.float64) / float(P - 1))[None, :]
    out = np.empty((n, P, 6), dtype=np.float64)
    for c in range(6):
        out[:, :, c] = _any(model, spread, seed, i, q, c, t, amp, pert)
    return out.reshape(n, 6 * P)


def default_pert(threshold, spline_points):
    """Perturbation amplitude that puts the median intra-cluster distance near the threshold,
    so about half of a cluster's pairs become edges (mean degree ~ cluster_size/2)."""
    P = spline_points
    s2 = sum((p / (P - 1.0)) ** 2 for p in range(P))
    return float(threshold) * 0.5 / (s2 ** 0.5)


# ---- device-side generation through the C ABI (include/scema_synth.h) into torch tensors --------
def device_offsets(seed, n, cluster_size, len_min, len_max, first=0):
    from . import binding
    off = np.empty(n + 1, dtype=np.uint64)
    rc = binding.lib().scema_synth_offsets(seed, first, n, cluster_size, len_min, len_max, off.ctypes.data)
    if rc:
        raise binding.ScemaError(rc, "synth_offsets")
    return off


def device_histories(seed, n, cluster_size, amp, pert, off, first=0, device="cuda", model=0, spread=0.0):
    """-> torch float64 tensor [off[n], 6] generated on the device (bit-identical to histories())."""
    import torch
    from . import binding
    d_off = torch.from_numpy(off.astype(np.int64)).to(device)
    steps = torch.empty((int(off[-1]), 6), dtype=torch.float64, device=device)
    rc = binding.lib().scema_synth_histories_model_device(model, spread, seed, first, n, cluster_size, amp, pert, d_off.data_ptr(),
                                                          steps.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc:
        raise binding.ScemaError(rc, "synth_histories_device")
    torch.cuda.current_stream().synchronize()
    return steps


def device_rows(seed, n, cluster_size, spline_points, amp, pert, first=0, device="cuda", model=0, spread=0.0):
    import torch
    from . import binding
    rows_t = torch.empty((n, 6 * spline_points), dtype=torch.float64, device=device)
    rc = binding.lib().scema_synth_rows_model_device(model, spread, seed, first, n, cluster_size, spline_points, amp, pert,
                                                     rows_t.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc:
        raise binding.ScemaError(rc, "synth_rows_device")
    torch.cuda.current_stream().synchronize()
    return rows_t
