"""oracle/ -- TEST INFRASTRUCTURE ONLY.

Holds (a) `scpr_oracle.c`, a plain-C CPU restatement of the reference's v4 encode/decode path and
(b) the recipe that compiles the UNMODIFIED reference core into `oracle/_ref/libscpr_ref.so`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import or execute anything from this directory, and only as the checker or the CPU baseline:
the product (`screenpressor_b200/`) never does.
"""
