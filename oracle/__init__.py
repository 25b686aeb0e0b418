"""CPU oracle for the smafa hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under smafa_b200/ does.  See smafa_oracle.h for parity status.
"""
