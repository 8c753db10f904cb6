"""oracle/ - TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithms for the hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import anything
from here; the product package (speech_b200/) never does.
"""
