"""CPU oracle for the PDHG hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this package.  Parity is UNPINNED by the reference (it has no such path); see
oracle/pdhg_oracle.c for what pins it instead.
"""
