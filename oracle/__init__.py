"""oracle/ -- TEST INFRASTRUCTURE (the parity checker), NOT PRODUCT CODE.

numpy/ctypes front-end of ``oracle/oracle_c.c`` (the CPU restatement of the reference path) and of
``oracle/_ref/libisoext_ref.so`` (the reference's own CUDA sources behind a C shim, see
``oracle/ref_shim.cu``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline
legs may import this package; ``isoext_b200`` never does.
"""
from .cpu import (  # noqa: F401
    build_c,
    case_histogram,
    cells_dense,
    dual_contouring,
    get_intersection,
    mc_dense,
    mc_sparse,
    points_cells,
    points_dense,
)
