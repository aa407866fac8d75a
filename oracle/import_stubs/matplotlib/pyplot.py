"""Import-only stand-in for matplotlib.pyplot.  TEST INFRASTRUCTURE."""
