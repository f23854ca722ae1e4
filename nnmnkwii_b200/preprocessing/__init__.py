"""Pre-processing pieces on the DTW hot path (drop-in for the names the aligners need from
``nnmnkwii.preprocessing``)."""
import numpy as np


def trim_zeros_frames(x, eps=1e-7, trim="b"):
    """Remove leading and/or trailing zeros frames (nnmnkwii/preprocessing/generic.py:291-332).

    Host-side utility with the reference's exact semantics; the aligners compute the same trailing
    lengths on the device (C ABI ``nnk_trim_lengths``) without touching the host.
    """
    assert trim in {"f", "b", "fb"}
    T, D = x.shape
    s = np.sum(np.abs(x), axis=1)
    s[s < eps] = 0.0
    if trim == "f":
        return x[len(x) - len(np.trim_zeros(s, trim=trim)):]
    elif trim == "b":
        end = len(np.trim_zeros(s, trim=trim)) - len(x)
        if end == 0:
            return x
        else:
            return x[:end]
    elif trim == "fb":
        f = len(np.trim_zeros(s, trim="f"))
        b = len(np.trim_zeros(s, trim="b"))
        end = b - len(x)
        if end == 0:
            return x[len(x) - f:]
        else:
            return x[len(x) - f: end]


__all__ = ["trim_zeros_frames"]
