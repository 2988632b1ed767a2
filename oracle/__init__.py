"""CPU oracle for the acav100m hot path (k-means SGD step / assign, greedy-MI selection).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it, and there only as the checker (or as the timed CPU baseline) -- never as a fallback for the CUDA
path.  ``acav100m_b200`` never imports this package.

Parity status: **pinned against outputs of the reference itself** -- the reference ships no golden
vectors for this path, so ``oracle/gen_golden.py`` imports the unmodified
reference files from ``/root/reference`` (three documented shims for absent third-party modules),
runs them seeded on CPU and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every oracle function against those fixtures.
"""
