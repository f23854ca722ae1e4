"""nnmnkwii_b200 -- B200-native (sm_100a) implementation of nnmnkwii's two numeric hot paths.

Drop-in for ``nnmnkwii.paramgen``, ``nnmnkwii.autograd`` (MLPG part),
``nnmnkwii.preprocessing.alignment`` and ``nnmnkwii.metrics.melcd``: same Python signatures, the
arithmetic in hand-written CUDA kernels behind the C ABI of include/nnk_b200.h.

Every functional submodule imports ``nnmnkwii_b200._lib``, which raises ImportError if
libnnk_b200.so has not been built (``python -m nnmnkwii_b200.build``): there is no CPU fallback.
(The package root itself stays importable so that the build module can run.)
"""
__version__ = "0.1.0"
