"""CPU oracle package — TEST INFRASTRUCTURE ONLY.

Nothing under eda_b200/ imports this package.  Allowed importers: tests/,
__graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).
"""
