"""CPU oracle — TEST INFRASTRUCTURE ONLY.

Nothing under echoseal_b200/ may import this package.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
and only as the checker / CPU baseline (never as the thing measured or shipped)."""
