"""Headless no-op stand-in for matplotlib (the driver only plots; USMain.py:232-243)."""
__version__ = "0.0-prt-shim"
def use(*a, **k): return None
