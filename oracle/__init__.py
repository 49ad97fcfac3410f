"""CPU oracles — TEST INFRASTRUCTURE ONLY (see oracle/oracle_n.c, oracle/oracle_r.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package. The product (realsensetracker_b200/) never does.
"""
