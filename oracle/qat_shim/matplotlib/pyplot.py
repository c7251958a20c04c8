"""TEST INFRASTRUCTURE ONLY -- no-op stand-in for matplotlib.pyplot (see matplotlib/__init__.py)."""


def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
