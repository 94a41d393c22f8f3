"""CPU oracle for the LZ77 hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
the reported CPU baseline.  The product (``lz77_b200``) never imports it.
"""
from .binding import (  # noqa: F401
    Oracle,
    build_oracle,
    oracle,
    ref_binary,
    ref_run,
)
