"""ORACLE package (test infrastructure, not product code): CPU restatements of the reference's lap-time path.

`reference_port` -- level 1, scalar Python with the reference's own SciPy/numpy calls, pinned bit for bit to the
unmodified reference by tests/golden; `c_oracle` / `lap_oracle.c` -- level 2, plain C, batch capable.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package."""
