"""TEST INFRASTRUCTURE — parity oracle for the powspec hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``powspec_b200``) never does.
"""
from .oracle import (  # noqa: F401
    Oracle, OracleResult, load_oracle, have_ref, have_port, build_port, build_ref,
)
