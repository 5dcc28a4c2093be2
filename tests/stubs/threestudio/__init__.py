"""Minimal stand-in for the threestudio core (not vendored by the reference, README.md:41): just enough
surface for the reference's renderer/*.py and geometry/gaussian_base.py to import and run UNCHANGED in
the plumbing tests.  Test infrastructure only."""
__version__ = "0.2.3"
__modules__ = {}


def register(name):
    def deco(cls):
        __modules__[name] = cls
        return cls
    return deco


def find(name):
    return __modules__[name]


def info(*a, **k):
    pass


warn = debug = info
