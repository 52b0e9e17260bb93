"""Empty stub: the reference imports matplotlib.pyplot for drawing only (vrp_network.py:4)."""
