"""Import-only stand-in for SimpleITK (not installed): the reference's parameter-study modules import it at
module level but their file writers never call it.  TEST INFRASTRUCTURE (oracle/gen_golden_r2.py)."""
