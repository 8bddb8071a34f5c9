"""CPU oracle for the MadIPM hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under madipm_jl_b200/ may import this package; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do, and only as the checker.
"""
