"""TEST INFRASTRUCTURE -- no-op stand-in for matplotlib (SVIM_plot.py:1-5 imports it)."""


def use(*args, **kwargs):
    return None
