"""Import-only stand-in for natsort (not installed); natural sort as the reference's reader uses it.
TEST INFRASTRUCTURE (oracle/gen_golden_r2.py)."""
import re


def natsorted(seq, key=None):
    key = key or (lambda v: v)
    split = lambda s: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(key(s)))]
    return sorted(seq, key=split)
