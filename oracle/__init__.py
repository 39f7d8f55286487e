"""CPU oracle for the two hot paths — TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is imported by the product package (riffusion-hobby_b200/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as
the checker or the timed CPU baseline, never as the thing shipped.
"""
