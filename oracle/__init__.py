"""oracle/ -- TEST INFRASTRUCTURE ONLY (see oracle/contract_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package qtorch_b200 never does.
"""
