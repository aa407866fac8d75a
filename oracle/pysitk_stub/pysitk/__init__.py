"""Minimal stand-in for the `pysitk` package (not installed here) so the
unmodified reference can be imported by oracle/gen_golden.py.  Only the six
symbols the reference's solver path touches are provided."""
