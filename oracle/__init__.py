"""CPU oracle for the nnmnkwii MLPG + DTW hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  ``nnmnkwii_b200`` never does (tests/test_no_oracle_in_product.py
enforces it).  See oracle/nnk_oracle.c for the parity status of each family.
"""
from .nnk_oracle import *  # noqa: F401,F403
