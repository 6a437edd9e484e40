"""Entropy-coded latent bitstream: what `size(use_torchac=True)` measures in the reference
(wisp/models/grids/latent_grid.py:155-172, multi_latent_decoder.py:174-186).

Pinned against the restatement of latent_grid.py:160-169 in oracle/latent_oracle.symbol_stream
(tests/test_host_logic.py::test_symbol_stream_and_cdf_match_the_oracle, tests/test_oracle_golden.py::
test_symbol_stream_round_trip): the int16 symbol stream (dense ranks), the per-channel histogram and the
float32 CDF table handed to the coder.
NOT pinned: the coded bytes themselves -- `torchac` is absent, unvendored and unpinned, and the
reference never decodes its stream. The coder here (csrc/arith_coder.inl) is verified by
decode round trip and by its length against the empirical entropy.
"""
import numpy as np
import torch

from . import _lib


def float_cdf(counts):
    """The reference's float32 CDF row: cat(0, cumsum(counts / counts.sum())) / last
    (latent_grid.py:166-169)."""
    cdf = torch.cumsum(counts / counts.sum(), dim=0)
    cdf = torch.cat((torch.zeros(1, dtype=cdf.dtype, device=cdf.device), cdf))
    return cdf / cdf[-1:]


def quantize_cdf(cdf_float):
    """float CDF [K+1] -> strictly increasing uint32 with 16-bit precision, total 65536:
    round(cdf * (2^16 - K)) + arange(K+1)."""
    K = cdf_float.numel() - 1
    if K + 1 > (1 << 16):
        raise ValueError("too many distinct symbols for a 16-bit CDF")
    q = torch.round(cdf_float.double().cpu() * float((1 << 16) - K)).to(torch.int64) + torch.arange(K + 1)
    return q.numpy().astype(np.uint32)


def integer_cdf(counts):
    """The coding model of the stored container (codec.py): the 16-bit CDF from the INTEGER histogram in exact
    integer arithmetic, floor(cum * (2^16 - K) / total) + arange(K + 1). Writer and reader derive it from the same
    stored counts, so -- unlike the float32 cumsum of float_cdf, whose last bit depends on summation order, torch build
    and device -- it is identical everywhere and the arithmetic decoder cannot desynchronise."""
    c = np.asarray(counts.cpu() if isinstance(counts, torch.Tensor) else counts, dtype=np.int64)
    K = int(c.size)
    if K + 1 > (1 << 16):
        raise ValueError("too many distinct symbols for a 16-bit CDF")
    if K == 0 or (c <= 0).any():
        raise ValueError("integer_cdf needs positive counts")
    cum = np.concatenate((np.zeros(1, dtype=np.int64), np.cumsum(c)))
    q = (cum * ((1 << 16) - K)) // cum[-1] + np.arange(K + 1, dtype=np.int64)
    return q.astype(np.uint32)


def dense_ranks(column, unique_vals):
    """Rounded latents -> rank of each value among the sorted unique values, int16
    (the reference's `mapping[weight]`, latent_grid.py:161-165)."""
    q = torch.round(column).long()
    return torch.searchsorted(unique_vals, q).to(torch.int16)


def encode_column(column, unique_vals, counts, exact=False):
    """Returns (stream bytes, cdf uint32[K+1]). exact: the integer coding model of the container (integer_cdf);
    otherwise the reference's float32 CDF (what size(use_torchac=True) measures)."""
    cdf = integer_cdf(counts) if exact else quantize_cdf(float_cdf(counts))
    ranks = dense_ranks(column, unique_vals).cpu().numpy()
    return _lib.ac_encode(ranks, cdf), cdf


def decode_column(stream, cdf, n, unique_vals):
    ranks = _lib.ac_decode(stream, cdf, n)
    return unique_vals.cpu()[torch.from_numpy(ranks.astype(np.int64))]


def coded_bits_from_table(column, unique_vals, counts):
    stream, _ = encode_column(column, unique_vals, counts)
    return len(stream) * 8


def coded_bits(symbols):
    """Bits of an integer tensor coded with its own empirical distribution."""
    unique_vals, counts = torch.unique(symbols, return_counts=True)
    return coded_bits_from_table(symbols.float(), unique_vals, counts)
