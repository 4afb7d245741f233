"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the tCG hot path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the
product package `optimization_b200` never does.
"""
