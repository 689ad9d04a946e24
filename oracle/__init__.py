"""ORACLE — CPU restatement of the TextureMixer hot path.  TEST INFRASTRUCTURE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import anything from this package; the product
(`texturemixer_b200`) never does and fails loudly without its CUDA library.
"""
