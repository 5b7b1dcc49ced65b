"""Empty stand-in for matplotlib.pyplot (see package docstring)."""
