"""Minimal stand-in for the reference's FrozenDict `.config` (configuration_utils.py:55-85): attribute and item
access, `.get`, read-only.  The pipelines only read from it."""


class FrozenConfig(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        object.__setattr__(self, "_frozen", True)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        raise TypeError("config is read-only")

    def __setitem__(self, name, value):
        if getattr(self, "_frozen", False):
            raise TypeError("config is read-only")
        super().__setitem__(name, value)
