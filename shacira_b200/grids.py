"""`HashGrid` and SHACIRA's `LatentGrid`, host-side mirror of `wisp/models/grids/hash_grid.py`
and `wisp/models/grids/latent_grid.py`: same constructors, classmethods, attribute and
state-dict names (`codebook`, `codebook_lod_sizes`, `codebook_lod_first_idx`, `latent_dec.*`,
`prob_model.*`), so they drop in under the reference's `app/image` and `app/nerf`.

What differs is only where the work happens: `interpolate`, `ent_loss` and `size` call the
B200 C-ABI kernels (`shacira_b200._lib`) instead of decoding the whole table with
`torch.matmul` and launching one interpolation kernel per level.
"""
import math
import os
from typing import Any, Dict, List

import numpy as np
import torch
import torch.nn as nn

from . import bitstream, grid_ops
from ._lib import ShaciraError, ERR_UNSUPPORTED
from . import _lib
from .latent_decoders import (DecoderIdentity, HierarchicalLatentDecoder, LatentDecoder, MultiLatentDecoder)
from .prob_models import BitEstimator


class _NoBLAS:
    """Placeholder for the kaolin OctreeAS the reference builds in every grid constructor
    (hash_grid.py:59-60, latent_grid.py:69-70). kaolin is an un-vendored dependency and the
    ray-marching acceleration structure is outside the hot path (out of scope)."""

    def __init__(self, level):
        self.level = level
        self.max_level = level

    def _missing(self, *a, **k):
        raise NotImplementedError(
            "ray marching / tracing / queries need kaolin's OctreeAS, which is outside the latent hash-grid "
            "hot path; pass samples to grid.interpolate() directly")

    raymarch = raytrace = query = _missing


def _make_blas(level):
    try:  # use the real structure when the reference stack is importable
        from wisp.accelstructs import OctreeAS  # type: ignore
        import kaolin.ops.spc as spc_ops  # type: ignore
        blas = OctreeAS.make_dense(level=level)
        pts = spc_ops.unbatched_get_level_points(blas.points, blas.pyramid, level).clone()
        return blas, pts
    except Exception:
        return _NoBLAS(level), torch.zeros(0, 3, dtype=torch.short)


class BLASGrid(nn.Module):
    """wisp/models/grids/blas_grid.py:29-87, reduced to what the hash grids use."""

    def __init__(self, blas):
        super().__init__()
        self.blas = blas

    def raymarch(self, *args, **kwargs):
        return self.blas.raymarch(*args, **kwargs)

    def raytrace(self, *args, **kwargs):
        return self.blas.raytrace(*args, **kwargs)

    def query(self, *args, **kwargs):
        return self.blas.query(*args, **kwargs)

    def interpolate(self, coords, lod_idx):
        raise NotImplementedError

    def supported_blas(self):
        return set()

    def public_properties(self) -> Dict[str, Any]:
        return {"Acceleration Structure": self.blas}


def geometric_resolutions(min_grid_res, max_grid_res, num_lods):
    """latent_grid.py:280-281 / hash_grid.py:178-179 (Instant-NGP eq. 2-3)."""
    b = np.exp((np.log(max_grid_res) - np.log(min_grid_res)) / (num_lods - 1))
    return [int(1 + np.floor(min_grid_res * (b ** l))) for l in range(num_lods)]


class _HashGridBase(BLASGrid):
    def _alloc_levels(self, resolutions, resolution_dim, channels, draw):
        """Per-level tables concatenated into one Parameter (latent_grid.py:93-112)."""
        self.resolutions = resolutions
        self.resolution_dim = resolution_dim
        self.num_lods = len(resolutions)
        self.active_lods = [x for x in range(self.num_lods)]
        self.max_lod = self.num_lods - 1
        self.codebook_size = 2 ** self.codebook_bitwidth
        self.register_buffer("codebook_lod_sizes", torch.zeros(self.num_lods, dtype=torch.int32))
        self.register_buffer("codebook_lod_first_idx", torch.zeros(self.num_lods, dtype=torch.int32))
        tables, offset = [], 0
        for lod, res in enumerate(resolutions):
            rows = min(self.codebook_size, res ** resolution_dim)
            tables.append(draw(torch.zeros(rows, channels)))
            self.codebook_lod_sizes[lod] = rows
            self.codebook_lod_first_idx[lod] = offset
            offset += rows
        self.codebook = nn.Parameter(torch.cat(tables, dim=0))
        # host copies used to build kernel parameters without a device sync
        self._first_idx_host = tuple(int(v) for v in self.codebook_lod_first_idx.tolist())

    def _first_idx(self):
        fi = getattr(self, "_first_idx_host", None)
        if fi is None or len(fi) != self.num_lods:
            fi = tuple(int(v) for v in self.codebook_lod_first_idx.tolist())
            self._first_idx_host = fi
        return fi

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._first_idx_host = None

    @staticmethod
    def _flatten(coords):
        output_shape = coords.shape[:-1]
        if coords.ndim == 3:
            coords = coords.reshape(-1, coords.shape[-1])
        return coords, output_shape

    def _aggregate(self, feats, output_shape, lod_idx):
        if "RENDERING_FINAL" in os.environ:  # latent_grid.py:372-375
            mask = torch.zeros_like(feats)
            mask[:, :lod_idx * self.feature_dim] = 1
            feats = feats * mask
        if self.multiscale_type == "cat":
            return feats.reshape(*output_shape, feats.shape[-1])
        if self.multiscale_type == "sum":
            L = len(self.resolutions)
            return feats.reshape(*output_shape, L, feats.shape[-1] // L).sum(-2)
        raise NotImplementedError

    def raymarch(self, rays, raymarch_type, num_samples, level=None):
        return self.blas.raymarch(rays, raymarch_type=raymarch_type, num_samples=num_samples, level=self.blas_level)

    def supported_blas(self):
        return {type(self.blas)}


class HashGrid(_HashGridBase):
    """Instant-NGP style multi-resolution hash grid. wisp/models/grids/hash_grid.py:21-287."""

    def __init__(self, feature_dim: int, resolutions: List[int], multiscale_type: str = "sum",
                 resolution_dim: int = 3, feature_std: float = 0.0, feature_bias: float = 0.0,
                 codebook_bitwidth: int = 8, blas_level: int = 7):
        self.blas_level = blas_level
        blas, dense_points = _make_blas(blas_level)
        super().__init__(blas)
        self.dense_points = dense_points
        self.num_cells = self.dense_points.shape[0]
        self.occupancy = torch.zeros(self.num_cells)
        self.feature_dim = feature_dim
        self.multiscale_type = multiscale_type
        self.feature_std = feature_std
        self.feature_bias = feature_bias
        self.codebook_bitwidth = codebook_bitwidth
        self._alloc_levels(resolutions, resolution_dim, feature_dim,
                           lambda t: t + torch.randn_like(t) * self.feature_std)

    def size(self, use_torchac=False, use_prob_model=False):
        return 0.0, self.codebook.numel() * torch.finfo(self.codebook.dtype).bits

    @classmethod
    def from_octree(cls, feature_dim, base_lod=2, num_lods=1, multiscale_type="sum", resolution_dim=3,
                    feature_std=0.0, feature_bias=0.0, codebook_bitwidth=8, blas_level=7):
        resolutions = [2 ** (base_lod + x) for x in range(num_lods)]
        return cls(feature_dim=feature_dim, resolutions=resolutions, multiscale_type=multiscale_type,
                   feature_std=feature_std, feature_bias=feature_bias, codebook_bitwidth=codebook_bitwidth,
                   blas_level=blas_level, resolution_dim=resolution_dim)

    @classmethod
    def from_geometric(cls, feature_dim, num_lods, multiscale_type="sum", resolution_dim=3, feature_std=0.0,
                       feature_bias=0.0, codebook_bitwidth=8, min_grid_res=16, max_grid_res=None, blas_level=7):
        resolutions = geometric_resolutions(min_grid_res, max_grid_res, num_lods)
        return cls(feature_dim=feature_dim, resolutions=resolutions, multiscale_type=multiscale_type,
                   feature_std=feature_std, feature_bias=feature_bias, codebook_bitwidth=codebook_bitwidth,
                   blas_level=blas_level, resolution_dim=resolution_dim)

    @classmethod
    def from_resolutions(cls, feature_dim, resolutions, multiscale_type="sum", resolution_dim=3, feature_std=0.0,
                         feature_bias=0.0, codebook_bitwidth=8, blas_level=7):
        return cls(feature_dim=feature_dim, resolutions=resolutions, multiscale_type=multiscale_type,
                   feature_std=feature_std, feature_bias=feature_bias, codebook_bitwidth=codebook_bitwidth,
                   blas_level=blas_level, resolution_dim=resolution_dim)

    def freeze(self):
        self.codebook.requires_grad_(False)

    def interpolate(self, coords, lod_idx):
        """coords [batch, (num_samples,) 2|3] -> features (hash_grid.py:222-255)."""
        coords, output_shape = self._flatten(coords)
        feats = grid_ops.hashgrid_any(coords, self.codebook, self._first_idx(), self.resolutions,
                                      self.codebook_bitwidth)
        return self._aggregate(feats, output_shape, lod_idx)

    def name(self) -> str:
        return "Hash Grid"

    def public_properties(self) -> Dict[str, Any]:
        lods = None if not self.active_lods else f"{min(self.active_lods)} - {max(self.active_lods)}"
        return {**super().public_properties(), "Feature Dims": self.feature_dim, "Total LODs": self.max_lod,
                "Active feature LODs": lods, "Interpolation": "linear",
                "Multiscale aggregation": self.multiscale_type, "HashTable Size": f"2^{self.codebook_bitwidth}"}


class LatentGrid(_HashGridBase):
    """SHACIRA's quantized-latent hash grid. wisp/models/grids/latent_grid.py:23-415."""

    def __init__(self, feature_dim: int, latent_dim: int, resolutions: List[int], multiscale_type: str = "sum",
                 resolution_dim: int = 3, feature_std: float = 0.0, feature_bias: float = 0.0,
                 codebook_bitwidth: int = 8, blas_level: int = 7, init_grid: str = "normal",
                 conf_latent_decoder: Dict[str, Any] = {}, conf_entropy_reg: Dict[str, Any] = {}):
        self.blas_level = blas_level
        blas, dense_points = _make_blas(blas_level)
        super().__init__(blas)
        self.dense_points = dense_points
        self.num_cells = self.dense_points.shape[0]
        self.occupancy = torch.zeros(self.num_cells)
        self.feature_dim = feature_dim
        self.latent_dim = feature_dim if latent_dim == 0 else latent_dim
        self.multiscale_type = multiscale_type
        self.feature_std = feature_std
        self.feature_bias = feature_bias
        self.codebook_bitwidth = codebook_bitwidth

        def draw(t):
            if init_grid == "uniform":
                return t + (torch.rand_like(t) - 0.5) * 2 * self.feature_std
            if init_grid == "normal":
                return t + torch.randn_like(t) * self.feature_std
            return t

        self._alloc_levels(resolutions, resolution_dim, self.latent_dim, draw)
        self.latent_dec = self.setup_decoders(conf_latent_decoder)
        self.prob_model = None
        self.noise = None
        self.noise_on_device = True  # False reproduces the reference's CPU RNG stream (latent_grid.py:128)
        if conf_latent_decoder["ldecode_enabled"] and (conf_entropy_reg["entropy_reg"] > 0.0
                                                        or conf_entropy_reg["entropy_reg_end"] > 0.0):
            self.prob_model = BitEstimator(self.latent_dim, num_layers=conf_entropy_reg["num_prob_layers"])
            self.noise_freq = conf_entropy_reg["noise_freq"]
        self.last_level_bits = None

    # ---- decoders -----------------------------------------------------------------------
    def setup_decoders(self, decoder_cfg):
        if not decoder_cfg["ldecode_enabled"]:
            return DecoderIdentity()
        decoder_cfg["feature_dim"] = self.feature_dim
        decoder_cfg["latent_dim"] = self.latent_dim
        kind = decoder_cfg["ldecode_type"]
        if kind == "hierarchical":
            offsets = list(self._first_idx()) + [int(self.codebook.shape[0])]  # see Q5 note in the class
            return HierarchicalLatentDecoder(self.num_lods, offsets, decoder_cfg)
        if kind == "multi":
            decoder_cfg["num_entries"] = self.codebook.size(0)
            decoder = MultiLatentDecoder(**decoder_cfg)
            del decoder_cfg["num_entries"]
            return decoder
        if kind == "single":
            return LatentDecoder(**decoder_cfg)
        raise ValueError("unknown ldecode_type %r" % (kind,))

    # ---- bit-rate loss ---------------------------------------------------------------------
    def _draw_noise(self):
        shape = self.codebook.shape
        if self.noise_on_device:
            return torch.rand(shape, device=self.codebook.device, dtype=self.codebook.dtype) - 0.5
        return torch.rand(shape).to(self.codebook) - 0.5  # latent_grid.py:128,130

    def ent_loss(self, idx, is_val=False):
        """(bits / rows, bits) of the factorized density (latent_grid.py:122-136), fused on the GPU."""
        if self.prob_model is None:
            return 0.0, 0.0
        noise = self.noise
        if is_val:
            noise = None                      # round(w): no draw needed (the reference draws and discards it)
        elif self.noise_freq == 1:
            noise = self._draw_noise()
        elif idx % self.noise_freq == 0:
            self.noise = self._draw_noise()
            noise = self.noise
        total_bits = grid_ops.entropy_bits(self.codebook, None if is_val else noise,
                                           self.prob_model.packed_params(), self.prob_model.num_layers,
                                           self._first_idx())
        return total_bits / self.codebook.shape[0], total_bits

    # ---- storage size ----------------------------------------------------------------------
    def symbol_statistics(self):
        """Per channel: (sorted unique rounded values, counts), both int64 on the table's device.
        Equals torch.unique(round(codebook[:, c]).long(), return_counts=True) (latent_grid.py:142-143),
        computed with one min/max pass and one histogram pass."""
        w = self.codebook.detach()
        _, mm = _lib.quantize_symbols(w, want_symbols=False)
        mm = mm.cpu()
        lo = [int(v) for v in mm[:, 0]]
        bins = int((mm[:, 1] - mm[:, 0]).max().item()) + 1
        counts = _lib.symbol_histogram(w, lo, bins)
        out = []
        for c in range(w.shape[1]):
            nz = torch.nonzero(counts[c], as_tuple=False).squeeze(1)
            out.append((nz + lo[c], counts[c][nz]))
        return out

    def size(self, use_torchac=False, use_prob_model=False):
        """(decoder bits, latent bits) as in latent_grid.py:138-174. With use_torchac the latent
        bits are the length of this package's own range-coded stream (torchac is an absent,
        unpinned dependency of the reference -- see shacira_b200/bitstream.py)."""
        ldec_size = self.latent_dec.size(use_torchac)
        codebook_bits = 0
        stats = self.symbol_statistics()
        for dim, (unique_vals, counts) in enumerate(stats):
            if not use_prob_model:
                probs = counts / torch.sum(counts)
            else:
                assert self.prob_model is not None
                probs = self.prob_model(unique_vals + 0.5, single_channel=dim) - \
                    self.prob_model(unique_vals - 0.5, single_channel=dim)
            if not use_torchac:
                information_bits = torch.clamp(-1.0 * torch.log(probs + 1e-10) / np.log(2.0), 0, 1000)
                codebook_bits += torch.sum(information_bits * counts).item()
            else:
                codebook_bits += bitstream.coded_bits_from_table(self.codebook.detach()[:, dim], unique_vals, counts)
        return ldec_size, codebook_bits

    # ---- constructors ----------------------------------------------------------------------
    @classmethod
    def from_octree(cls, feature_dim, latent_dim=0, base_lod=2, num_lods=1, multiscale_type="sum", resolution_dim=3,
                    feature_std=0.0, feature_bias=0.0, codebook_bitwidth=8, blas_level=7, init_grid="normal",
                    conf_latent_decoder={}, conf_entropy_reg={}):
        resolutions = [2 ** (base_lod + x) for x in range(num_lods)]
        return cls(feature_dim=feature_dim, resolutions=resolutions, multiscale_type=multiscale_type,
                   feature_std=feature_std, feature_bias=feature_bias, codebook_bitwidth=codebook_bitwidth,
                   blas_level=blas_level, latent_dim=latent_dim, conf_latent_decoder=conf_latent_decoder,
                   conf_entropy_reg=conf_entropy_reg, resolution_dim=resolution_dim, init_grid=init_grid)

    @classmethod
    def from_geometric(cls, feature_dim, num_lods, latent_dim=0, multiscale_type="sum", resolution_dim=3,
                       feature_std=0.0, feature_bias=0.0, codebook_bitwidth=8, min_grid_res=16, max_grid_res=None,
                       blas_level=7, init_grid="normal", conf_latent_decoder={}, conf_entropy_reg={}):
        resolutions = geometric_resolutions(min_grid_res, max_grid_res, num_lods)
        return cls(feature_dim=feature_dim, resolutions=resolutions, multiscale_type=multiscale_type,
                   feature_std=feature_std, feature_bias=feature_bias, codebook_bitwidth=codebook_bitwidth,
                   blas_level=blas_level, latent_dim=latent_dim, conf_latent_decoder=conf_latent_decoder,
                   conf_entropy_reg=conf_entropy_reg, resolution_dim=resolution_dim, init_grid=init_grid)

    @classmethod
    def from_resolutions(cls, feature_dim, resolutions, latent_dim=0, multiscale_type="sum", resolution_dim=3,
                         feature_std=0.0, feature_bias=0.0, codebook_bitwidth=8, blas_level=7, init_grid="normal",
                         conf_latent_decoder={}, conf_entropy_reg={}):
        return cls(feature_dim=feature_dim, resolutions=resolutions, multiscale_type=multiscale_type,
                   feature_std=feature_std, feature_bias=feature_bias, codebook_bitwidth=codebook_bitwidth,
                   blas_level=blas_level, latent_dim=latent_dim, conf_latent_decoder=conf_latent_decoder,
                   conf_entropy_reg=conf_entropy_reg, resolution_dim=resolution_dim, init_grid=init_grid)

    def freeze(self):
        self.codebook.requires_grad_(False)
        for p in self.latent_dec.parameters():
            p.requires_grad_(False)
        if self.prob_model is not None:
            for p in self.prob_model.parameters():
                p.requires_grad_(False)

    # ---- the hot path ------------------------------------------------------------------------
    def interpolate(self, coords, lod_idx):
        """coords [batch, (num_samples,) 2|3] -> decoded multi-level features
        (latent_grid.py:340-382). Affine decoders run fused in one kernel; anything else is decoded
        table-side on the GPU and interpolated by the plain kernel."""
        coords, output_shape = self._flatten(coords)
        dec = self.latent_dec
        fused = isinstance(dec, (LatentDecoder, HierarchicalLatentDecoder)) and dec.is_affine()
        if fused and self.latent_dim in (1, 2, 4) and self.feature_dim in (1, 2, 4, 8):
            A, shift = dec.affine_map()
            if dec.use_sga:
                # SGA mixes floor/ceil with Gumbel noise: RNG-bound table-side pre-pass
                latents, round_flag = dec.quantize(self.codebook), False
            else:
                latents, round_flag = self.codebook, True  # rounding + straight-through inside the kernel
            feats = grid_ops.latent_hashgrid(coords, latents, A, shift, self._first_idx(), self.resolutions,
                                             self.codebook_bitwidth, round_flag)
        else:
            table = dec(self.codebook)  # identity / multi / non-affine decoders
            if table.shape[1] not in (1, 2, 4, 8):
                raise ShaciraError(ERR_UNSUPPORTED, "decoded feature_dim %d has no compiled kernel" % table.shape[1])
            feats = grid_ops.hashgrid_any(coords, table, self._first_idx(), self.resolutions, self.codebook_bitwidth)
        return self._aggregate(feats, output_shape, lod_idx)

    def name(self) -> str:
        return "Latent Grid"

    def public_properties(self) -> Dict[str, Any]:
        lods = None if not self.active_lods else f"{min(self.active_lods)} - {max(self.active_lods)}"
        return {**super().public_properties(), "Feature Dims": self.feature_dim, "Latent Dims": self.latent_dim,
                "Total LODs": self.max_lod, "Active feature LODs": lods, "Interpolation": "linear",
                "Multiscale aggregation": self.multiscale_type, "HashTable Size": f"2^{self.codebook_bitwidth}"}
