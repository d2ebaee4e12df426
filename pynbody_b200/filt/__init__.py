"""Tree-less geometric selections of the reference's ``pynbody.filt`` that reach native code."""
from . import geometry_selection

__all__ = ["geometry_selection"]
