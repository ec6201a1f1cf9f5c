"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle/morpho_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  Nothing under morpho_b200/ does.
"""
