"""gridfluidsim3d_b200 -- B200-native PIC/FLIP particle<->grid transfer path for GridFluidSim3D.

The product is the CUDA library behind include/gfs_b200.h (csrc/, built into libgfs_b200.so); this
package is the thin host-side mirror used by the tests and bench.py.  Nothing here falls back to a CPU
implementation.
"""
from . import capi, synth  # noqa: F401
from .capi import Context, GfsError  # noqa: F401

__all__ = ["capi", "synth", "Context", "GfsError"]
