"""Import-only stand-in for matplotlib (not installed).  TEST INFRASTRUCTURE (oracle/gen_golden_r2.py)."""
