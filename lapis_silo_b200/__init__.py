"""lapis_silo_b200 — B200 (sm_100a) drop-in for the bitmap filter + Mutations hot path of RhyDB/SILO.

The product is `libsilo_b200.so` (csrc/, C ABI in include/silo_b200.h) plus the C++ host layer that
mirrors the reference's operator interface (host/). This Python package is only the harness that
tests and bench.py use to reach them; it contains no compute path and no CPU fallback.
"""
