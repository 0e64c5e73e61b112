"""laris_b200 - B200-native drop-in for LARIS's per-cell spatial LR scoring path.

Mirrors the reference's aliases (reference laris/__init__.py:38-47):
``laris_b200.tl`` (tools) and ``laris_b200.pp`` (preprocessing).  Plotting is
out of scope.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy so that ``import laris_b200.synth`` does not need the CUDA library
    import importlib
    if name in ("tl", "tools"):
        return importlib.import_module(".tools", __name__)
    if name in ("pp", "preprocessing"):
        return importlib.import_module(".preprocessing", __name__)
    if name in ("dist", "engine", "synth"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
