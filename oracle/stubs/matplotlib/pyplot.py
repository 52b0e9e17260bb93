"""Empty stub of matplotlib.pyplot (drawing is out of scope)."""
