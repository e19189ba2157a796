"""oracle/ -- TEST INFRASTRUCTURE.  CPU restatement + compiled reference used as the parity
checker.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may
import this package; the product (julius_b200/) never does."""
