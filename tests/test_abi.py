"""CPU: the C-ABI library builds, loads, exports every symbol include/nnk_b200.h declares, and the
product never routes through the oracle or any CPU fallback.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header():
    return open(os.path.join(ROOT, "include", "nnk_b200.h")).read()


def test_library_exports_every_declared_symbol():
    from nnmnkwii_b200 import _lib
    declared = sorted(set(re.findall(r"\b(nnk_[a-z0-9_]+)\s*\(", _header())))
    assert len(declared) >= 10
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "libnnk_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared
    assert _lib.lib.nnk_abi_version() == int(re.search(r"#define NNK_ABI_VERSION (\d+)", _header()).group(1))


def test_struct_layouts_match_header():
    from nnmnkwii_b200 import _lib
    h = _header()
    assert int(re.search(r"#define NNK_MAX_WIN (\d+)", h).group(1)) == _lib.NNK_MAX_WIN
    assert int(re.search(r"#define NNK_MAX_HALF (\d+)", h).group(1)) == _lib.NNK_MAX_HALF
    assert ctypes.sizeof(_lib.NnkWindows) == 4 + 4 * _lib.NNK_MAX_WIN * 2 + 4 + 8 * _lib.NNK_MAX_WIN * _lib.NNK_MAX_TAPS
    assert ctypes.sizeof(_lib.NnkStatus) == 16
    assert _lib.CHAIN_DTYPE.itemsize == 16
    # field order of nnk_mlpg_args_t
    body = re.search(r"typedef struct nnk_mlpg_args \{(.*?)\} nnk_mlpg_args_t;", h, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    assert names == [f[0] for f in _lib.NnkMlpgArgs._fields_]


def test_status_decode_roundtrip():
    from nnmnkwii_b200 import _lib
    st = _lib.NnkStatus()
    _lib.lib.nnk_status_decode(ctypes.c_uint64(0), ctypes.byref(st))
    assert st.code == 0
    key = (7 << 42) | (5 << 21) | 123
    _lib.lib.nnk_status_decode(ctypes.c_uint64((~key) & 0xFFFFFFFFFFFFFFFF), ctypes.byref(st))
    assert (st.code, st.utt, st.chain, st.frame) == (1, 7, 5, 123)


def test_window_validation_matches_reference_asserts():
    import numpy as np
    from nnmnkwii_b200 import _lib
    with pytest.raises(AssertionError):  # len(coeff) != l + u + 1, paramgen/_mlpg.py:45
        _lib.make_windows([(1, 1, np.array([1.0, 2.0]))])
    with pytest.raises(NotImplementedError):
        _lib.make_windows([(0, 0, np.array([1.0]))] * 9)


def test_product_never_touches_the_oracle_or_a_cpu_fallback():
    pkg = os.path.join(ROOT, "nnmnkwii_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "nnk_oracle" not in src and "oracle/_ref" not in src, f
                assert "/root/reference" not in src, f


def test_cuda_sources_target_sm100a_only():
    from nnmnkwii_b200 import build
    assert "arch=compute_100a,code=sm_100a" in build.NVCC_FLAGS
    assert not any("sm_90" in f or "sm_80" in f for f in build.NVCC_FLAGS)


def test_no_gpu_means_loud_failure():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nnmnkwii_b200 import paramgen as G
    with pytest.raises(Exception) as ei:
        G.mlpg(np.zeros((4, 3), np.float32), np.ones((4, 3), np.float32), [(0, 0, np.array([1.0]))])
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_host_helpers_match_reference_semantics(golden):
    import numpy as np
    from conftest import windows_set
    from nnmnkwii_b200 import paramgen as G
    # reshape_means: tests/test_paramgen.py:98-110
    m = np.random.rand(7, 6)
    r = G.reshape_means(m, 2)
    assert r.shape == (21, 2) and np.array_equal(G.reshape_means(r, 2), r)
    assert np.array_equal(r, m.reshape(7, 3, 2).transpose(1, 0, 2).reshape(-1, 2))
    # build_win_mats / full_window_mat: tests/test_paramgen.py:62-79
    for ws in windows_set():
        wm = G.build_win_mats(ws, 6)
        full = G.full_window_mat(wm, 6)
        assert full.shape == (6 * len(ws), 6)
        for i, (l, u, c) in enumerate(ws):
            blk = full[6 * i:6 * (i + 1)]
            for t in range(6):
                for k in range(-l, u + 1):
                    if 0 <= t + k < 6:
                        assert blk[t, t + k] == c[l + k]
            assert wm[i].l == l and wm[i].u == u and wm[i].transposed
