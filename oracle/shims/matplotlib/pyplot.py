"""TEST INFRASTRUCTURE -- no-op pyplot: every attribute is a callable that returns a dummy."""


class _Dummy(object):
    def __getattr__(self, name):
        return _Dummy()

    def __call__(self, *args, **kwargs):
        return _Dummy()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Dummy()
