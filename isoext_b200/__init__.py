"""isoext_b200 -- Blackwell-native iso-surface extraction behind the isoext Python API.

Same public names as the reference package (src/isoext/__init__.py:6-26).  The extension names
(``UniformGrid`` ... ``marching_cubes``) are backed by hand-written sm_100a CUDA kernels reached
through the C-ABI library ``libisoext_b200.so`` (include/isoext_b200.h); there is no CPU path.
"""
import importlib.util

if importlib.util.find_spec("torch") is None:  # same guard as the reference (src/isoext/__init__.py:3-4)
    raise ImportError("PyTorch is required but not installed. Please install PyTorch with CUDA support.\n")

from . import sdf, utils  # noqa: E402,F401
from .dc import Intersection, dual_contouring, get_intersection  # noqa: E402,F401
from .grid import Grid, ImplicitGrid, UniformGrid  # noqa: E402,F401
from .mc import marching_cubes  # noqa: E402,F401
from .sparse import SparseGrid  # noqa: E402,F401
from .utils import gaussian_smooth, make_grid, write_obj, write_ply  # noqa: E402,F401

# the reference's names, plus the extensions ImplicitGrid (analytic SDF evaluated in the kernels) and write_ply
__all__ = [
    "ImplicitGrid",
    "Intersection",
    "SparseGrid",
    "UniformGrid",
    "dual_contouring",
    "get_intersection",
    "gaussian_smooth",
    "marching_cubes",
    "make_grid",
    "write_obj",
    "write_ply",
]
