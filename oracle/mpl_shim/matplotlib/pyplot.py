"""See matplotlib/__init__.py -- attribute access returns a do-nothing callable."""


def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
