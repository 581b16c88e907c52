"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference algorithms used as parity checkers.  Nothing in
``monopsr_b200`` may import this package; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.
"""
