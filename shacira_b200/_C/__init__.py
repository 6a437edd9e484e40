"""Drop-in for the reference's pybind11 module `wisp._C` (wisp/csrc/bindings.cpp:18-28).

Only the `ops` submodule exists: `render` (find_depth_bound) and `external` (mesh2sdf) are
outside the hot path. `shacira_b200.compat.install_as_wisp_C()` registers
this package as `wisp._C` so the reference's `wisp/ops/grid.py` runs unmodified on it.
"""
from . import ops  # noqa: F401
