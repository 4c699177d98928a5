"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference algorithm for the L4P inference hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
this package, and only as the checker / the CPU baseline. The product (`l4p_b200/`) never imports it.
"""
