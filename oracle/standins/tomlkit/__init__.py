"""Stand-in for ``tomlkit`` (TEST SCAFFOLDING; see ../xarray)."""


def dumps(obj):
    return str(obj)
