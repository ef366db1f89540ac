"""On-disk formats either side of the path.

* pulse files of the reference's examples: whitespace-separated I/Q columns in Hz, one row per time slice, scaled by
  1e-9 to GHz on load (examples/cavity_qubit.jl:16-18, examples/zz_coupling_simulation.jl:3-4);
* compress_states / decompress_states (src/utils.jl:96-109): pack the columns of two state groups that live on
  disjoint row sets into max(n1, n2) columns.
"""
from __future__ import annotations

import numpy as np

__all__ = ["read_pulse", "write_pulse", "compress_states", "decompress_states"]


def read_pulse(path, scale=1e-9):
    """-> u (2 x Nt): row 0 = I, row 1 = Q, in GHz (file in Hz)."""
    iq = np.loadtxt(path, ndmin=2)
    if iq.shape[1] != 2:
        raise ValueError(f"{path}: expected two whitespace-separated columns (I, Q), got {iq.shape[1]}")
    return np.ascontiguousarray(iq.T * scale)


def write_pulse(path, u, scale=1e-9):
    """Inverse of read_pulse: u (2 x Nt, GHz) -> I/Q columns in Hz with full double precision."""
    u = np.asarray(u, dtype=np.float64)
    if u.ndim != 2 or u.shape[0] != 2:
        raise ValueError("u must be 2 x Nt (I, Q)")
    np.savetxt(path, (u / scale).T, fmt="%.17g")


def compress_states(x, v):
    """src/utils.jl:96-102.  v = ((rows1, cols1), (rows2, cols2)) with 0-based index arrays."""
    (r1, c1), (r2, c2) = v
    n1, n2 = len(c1), len(c2)
    x = np.asarray(x)
    out = np.zeros((x.shape[0], max(n1, n2)), dtype=x.dtype)
    out[np.ix_(r1, range(n1))] = x[np.ix_(r1, c1)]
    out[np.ix_(r2, range(n2))] = x[np.ix_(r2, c2)]
    return out


def decompress_states(x_compr, v):
    """src/utils.jl:103-109."""
    (r1, c1), (r2, c2) = v
    n1, n2 = len(c1), len(c2)
    x_compr = np.asarray(x_compr)
    out = np.zeros((x_compr.shape[0], n1 + n2), dtype=x_compr.dtype)
    out[np.ix_(r1, c1)] = x_compr[np.ix_(r1, range(n1))]
    out[np.ix_(r2, c2)] = x_compr[np.ix_(r2, range(n2))]
    return out
