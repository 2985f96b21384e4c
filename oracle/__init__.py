"""CPU oracle for the cnf_ot flow train step (test infrastructure, not product).

Only `tests/`, `__graft_entry__.smoke()` and bench.py's `cpu_baseline` /
`--impl reference` legs may import this package.  See `oracle/rqs.py`.
"""
