"""shacira_b200 -- B200-native (sm_100a) latent multi-resolution hash-grid hot path of SHACIRA.

Public surface (mirrors the reference's wisp modules, see DESIGN.md / INTEGRATION.md):
    shacira_b200.grid_ops          <- wisp/ops/grid.py           hashgrid, hashgrid2d, latent_hashgrid
    shacira_b200._C.ops            <- wisp._C.ops                4 pybind entry points
    shacira_b200.grids             <- wisp/models/grids          HashGrid, LatentGrid
    shacira_b200.latent_decoders   <- wisp/models/latent_decoders
    shacira_b200.prob_models       <- wisp/models/prob_models
All compute goes through libshacira_b200.so (include/shacira_b200.h); there is no fallback.
"""
__version__ = "0.1.0"
