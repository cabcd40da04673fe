"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference algorithms.

Nothing under oracle/ is imported by the product package `starst3r_b200`; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may use it, and only as the checker / the timed CPU baseline.
"""
