"""On-disk format of a fitted model: the artifact whose size the reference only COUNTS.

The reference reports `BPP = (latent bits + 32 * #latent-decoder params + 32 * #non-grid params) / (H * W)`
(wisp/trainers/image_trainer.py:162-168, 471-502) where the latent bits are `8 * len(torchac byte stream)` of the
rounded latents coded with their own histogram (wisp/models/grids/latent_grid.py:155-172) -- but it never writes that
stream anywhere, never decodes it, and pickles whole pipelines instead (image_trainer.py:477-483). This module makes
the number a verifiable file:

    blob = codec.encode_model(grid, mlp)          # bytes; len(blob) * 8 / (H * W) is a true bpp
    state = codec.decode_model(blob)              # integer latents, decoder parameters, MLP, grid geometry
    codec.load_into(state, grid2, mlp2)           # a freshly constructed LatentGrid / MLP renders the same image

Layout (little endian):
    magic "SHCR" | u16 version | u32 header bytes | header (JSON: geometry, decoder config, section table)
    per latent channel: i32 first symbol | u32 K | u32 counts[K] | u64 stream bytes | arithmetic-coded dense ranks
    float32 sections in the order of the header's table: latent_dec.div / scale / shift, every MLP tensor

The per-channel histogram IS the coding model: decoder and encoder rebuild the identical 16-bit CDF from the counts
(bitstream.float_cdf -> quantize_cdf, the float32 CDF being the reference's own, latent_grid.py:166-169). What the
reference's formula leaves out -- the histogram itself and the header -- is reported separately by `size_report`.
Decoding restores round(latent) exactly; that is all the decoder path reads (straight-through rounding, the
`ŵ = round(w)` of basic_latent_decoder.py:182-198), so the decoded model renders bit-identically to the fitted one.
Entropy coding is host-side byte work (csrc/arith_coder.inl through the C ABI), and so is its coding model: the
histogram is taken on the host from the very symbols that are coded.
"""
import json
import struct
import zlib

import numpy as np
import torch

from . import _lib, bitstream

MAGIC = b"SHCR"
VERSION = 2   # 2: integer coding model + CRC32 of every channel's symbols


def _channel_stats(grid):
    """Per channel (sorted unique rounded values, counts) as int64 CPU tensors. The symbols have to reach the host
    for the coder anyway, so the coding model is always built there, from the same bytes that get coded (one path,
    whatever device the table lives on)."""
    q = torch.round(grid.codebook.detach()).to(torch.int64).cpu()
    return [torch.unique(q[:, c], return_counts=True) for c in range(q.shape[1])]


def _f32_bytes(t):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype("<f4")).tobytes()


def encode_model(grid, mlp=None, extra_meta=None):
    """Serialise a LatentGrid (integer latents entropy coded, latent decoder in fp32) and optionally the decoder MLP."""
    w = grid.codebook.detach()
    T, C = w.shape
    dec = grid.latent_dec
    sections, payload = [], []

    def add(name, t):
        sections.append({"name": name, "shape": list(t.shape)})
        payload.append(_f32_bytes(t))

    for name, p in dec.state_dict().items():
        add("latent_dec." + name, p)
    if mlp is not None:
        for name, p in mlp.state_dict().items():
            add("mlp." + name, p)
    header = {
        "rows": int(T), "latent_dim": int(C), "feature_dim": int(grid.feature_dim), "num_lods": int(grid.num_lods),
        "resolutions": [int(r) for r in grid.resolutions], "codebook_bitwidth": int(grid.codebook_bitwidth),
        "resolution_dim": int(grid.resolution_dim),
        "multiscale_type": str(grid.multiscale_type), "sections": sections, "meta": extra_meta or {},
    }
    hbytes = json.dumps(header, separators=(",", ":")).encode("utf-8")
    out = [MAGIC, struct.pack("<HI", VERSION, len(hbytes)), hbytes]
    for c, (uniq, counts) in enumerate(_channel_stats(grid)):
        lo, hi = int(uniq[0]), int(uniq[-1])
        K = hi - lo + 1
        if K > (1 << 15):
            raise ValueError("latent range %d..%d too wide for int16 symbols" % (lo, hi))
        # the reference's coding model: dense ranks over the OBSERVED values (latent_grid.py:161-165)
        # coding model from the integer counts in exact integer arithmetic: the reader rebuilds the identical CDF
        stream, _ = bitstream.encode_column(w[:, c].cpu(), uniq, counts, exact=True)
        dense = np.zeros(K, dtype="<u4")
        dense[(uniq - lo).numpy()] = counts.numpy().astype("<u4")
        crc = zlib.crc32(torch.round(w[:, c]).to(torch.int32).cpu().numpy().astype("<i4").tobytes()) & 0xFFFFFFFF
        out += [struct.pack("<iI", lo, K), dense.tobytes(), struct.pack("<QI", len(stream), crc), stream]
    out += payload
    return b"".join(out)


def decode_model(blob):
    """Inverse of encode_model: dict(header=..., latents=int64 [rows, C], tensors={name: float32 tensor})."""
    if blob[:4] != MAGIC:
        raise ValueError("not a SHCR stream")
    version, hlen = struct.unpack_from("<HI", blob, 4)
    if version != VERSION:
        raise ValueError("unsupported SHCR version %d" % version)
    pos = 10
    header = json.loads(blob[pos:pos + hlen].decode("utf-8"))
    pos += hlen
    T, C = header["rows"], header["latent_dim"]
    cols = []
    for _ in range(C):
        lo, K = struct.unpack_from("<iI", blob, pos)
        pos += 8
        dense = np.frombuffer(blob, dtype="<u4", count=K, offset=pos).astype(np.int64)
        pos += 4 * K
        slen, crc = struct.unpack_from("<QI", blob, pos)
        pos += 12
        nz = np.nonzero(dense)[0]
        if int(dense.sum()) != T:
            raise ValueError("SHCR stream: histogram of channel %d does not sum to the row count" % len(cols))
        uniq = torch.from_numpy(nz + lo)
        counts = torch.from_numpy(dense[nz])
        cdf = bitstream.integer_cdf(counts)
        col = bitstream.decode_column(blob[pos:pos + slen], cdf, T, uniq)
        if (zlib.crc32(col.to(torch.int32).numpy().astype("<i4").tobytes()) & 0xFFFFFFFF) != crc:
            raise ValueError("SHCR stream: checksum mismatch in channel %d (corrupt stream or coding model)" % len(cols))
        cols.append(col)
        pos += slen
    tensors = {}
    for sec in header["sections"]:
        n = int(np.prod(sec["shape"])) if sec["shape"] else 1
        arr = np.frombuffer(blob, dtype="<f4", count=n, offset=pos).reshape(sec["shape"]).copy()
        tensors[sec["name"]] = torch.from_numpy(arr)
        pos += 4 * n
    if pos != len(blob):
        raise ValueError("trailing bytes in SHCR stream")
    return {"header": header, "latents": torch.stack(cols, dim=1), "tensors": tensors}


def load_into(state, grid, mlp=None):
    """Fill a freshly constructed LatentGrid (same geometry / decoder config) and MLP from a decoded stream."""
    h = state["header"]
    if tuple(grid.codebook.shape) != (h["rows"], h["latent_dim"]) or \
            [int(r) for r in grid.resolutions] != h["resolutions"] or int(grid.codebook_bitwidth) != h["codebook_bitwidth"]:
        raise ValueError("grid geometry does not match the stream")
    with torch.no_grad():
        grid.codebook.copy_(state["latents"].to(grid.codebook))
        dec_sd = {k[len("latent_dec."):]: v for k, v in state["tensors"].items() if k.startswith("latent_dec.")}
        grid.latent_dec.load_state_dict({k: v.to(grid.codebook.device) for k, v in dec_sd.items()})
        if mlp is not None:
            mlp_sd = {k[len("mlp."):]: v for k, v in state["tensors"].items() if k.startswith("mlp.")}
            mlp.load_state_dict(mlp_sd)
    return grid, mlp


def size_report(grid, mlp, blob, pixels):
    """The reference's BPP terms next to the real file: what its formula counts, and what it leaves out."""
    dec_bits = 32 * sum(p.numel() for p in grid.latent_dec.parameters())
    mlp_bits = 32 * sum(p.numel() for p in mlp.parameters()) if mlp is not None else 0
    entropy_bits = 0.0
    stream_bits = 0
    hist_bits = 0
    for c, (uniq, counts) in enumerate(_channel_stats(grid)):
        p = counts.double() / counts.sum()
        entropy_bits += float((torch.clamp(-torch.log2(p + 1e-10), 0, 1000) * counts).sum())   # latent_grid.py:150-153
        stream_bits += bitstream.coded_bits_from_table(grid.codebook.detach()[:, c].cpu(), uniq, counts)
        hist_bits += 32 * (int(uniq[-1]) - int(uniq[0]) + 1) + 64 + 64 + 32
    ref_formula_bits = stream_bits + dec_bits + mlp_bits
    return {
        "file_bytes": len(blob), "file_bpp": len(blob) * 8.0 / pixels,
        "reference_formula_bpp": ref_formula_bits / pixels, "empirical_entropy_bpp": (entropy_bits + dec_bits + mlp_bits) / pixels,
        "latent_stream_bits": stream_bits, "latent_entropy_bits": entropy_bits,
        "uncounted_by_reference_bits": len(blob) * 8 - ref_formula_bits, "histogram_bits": hist_bits,
    }
