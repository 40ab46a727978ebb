"""Pass-through stand-in for ``func_timeout`` (TEST SCAFFOLDING; see ../xarray)."""


class FunctionTimedOut(Exception):
    pass


def func_timeout(timeout, func, args=(), kwargs=None):
    return func(*args, **(kwargs or {}))
