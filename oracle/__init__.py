"""CPU oracle for the TN.contract() hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this package.  The product (contractn_b200) never does.
"""
