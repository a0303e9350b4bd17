"""Process-wide engine handle.

`get_engine()` creates the CUDA engine on first use and raises if the CUDA library or a
device is missing -- there is no CPU fallback in this package.  Tests that exercise the host
logic without a GPU inject an engine bound to the kernel-source emulator (tests/emul) through
`set_engine`; the package itself never constructs anything but the CUDA engine.
"""
from .engine import SubdivisionEngine

_engine = None


def get_engine():
    global _engine
    if _engine is None:
        _engine = SubdivisionEngine()
    return _engine


def set_engine(engine):
    global _engine
    _engine = engine
    return engine
