"""Drop-in hooks for running the reference's own apps on this package.

    import shacira_b200.compat as compat
    compat.install_as_wisp_C()      # wisp.ops.grid now launches the B200 kernels
    compat.install_grids()          # wisp.models.grids.{HashGrid,LatentGrid} -> fused implementations
"""
import sys


def install_as_wisp_C():
    """Register shacira_b200._C as `wisp._C` (the reference imports it in wisp/ops/grid.py:10)."""
    from . import _C
    sys.modules["wisp._C"] = _C
    sys.modules["wisp._C.ops"] = _C.ops
    wisp = sys.modules.get("wisp")
    if wisp is not None:
        wisp._C = _C
    return _C


def install_grids():
    """Swap the reference's grid classes for the fused ones in an imported `wisp`."""
    import wisp.models.grids as wg  # type: ignore
    from . import grids
    wg.HashGrid = grids.HashGrid
    wg.LatentGrid = grids.LatentGrid
    for mod in ("wisp.models.grids.hash_grid", "wisp.models.grids.latent_grid"):
        m = sys.modules.get(mod)
        if m is not None:
            if hasattr(m, "HashGrid"):
                m.HashGrid = grids.HashGrid
            if hasattr(m, "LatentGrid"):
                m.LatentGrid = grids.LatentGrid
    return grids
