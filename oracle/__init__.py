"""CPU oracle for the EditAnything hot path — TEST INFRASTRUCTURE ONLY.

Nothing under editanything_b200/ may import this package; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do, and only as the checker / baseline.
"""
