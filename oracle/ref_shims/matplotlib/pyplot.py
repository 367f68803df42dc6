"""No-op stand-in for matplotlib.pyplot (test infrastructure only)."""


def quiver(*args, **kwargs):
    pass


def show(*args, **kwargs):
    pass
