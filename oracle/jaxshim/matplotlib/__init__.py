"""Empty stub: pendulum_sys.py:7 imports matplotlib.pyplot at module top; never used on the path."""
