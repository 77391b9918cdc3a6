"""Host-side replica of the kernels' counter-based dropout RNG (csrc/kernels/sm100_ptx.cuh:
``dropout_seed`` / ``dropout_keep4``): regenerates the exact keep mask of any dropout site from
``{seed, step}`` (the device-resident ``RngState``), the site's stream id and the element index.

Used by the tests to feed the fp32 oracle the SAME masks the native kernels applied (exact
dropout parity instead of a statistical check), and handy for debugging a diverging run.

Element indexing of the sites (flat index i; 4 consecutive elements share one 64-bit draw):
  * GEMM / LayerNorm-backward / embedding sites:  i = row * N + col       of the [M, N] tensor
  * attention probabilities:                      i = ((b * heads + h) * S + q) * S + k
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):           # uint64 arithmetic wraps, exactly like the device
        x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def site_seed(seed: int, step: int, stream: int) -> np.uint64:
    with np.errstate(over="ignore"):
        s = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(0x632BE59BD9B4E019) * np.uint64(step + 1)
        return _splitmix64(np.asarray(s, dtype=np.uint64))[()] ^ (np.uint64(stream) << np.uint64(32))


def keep_mask(seed: int, step: int, stream: int, shape, p: float) -> np.ndarray:
    """Boolean keep mask (True = kept, scaled by 1 / (1 - p) by the kernels) of a dropout site
    whose elements are indexed row-major over ``shape``."""
    n = int(np.prod(shape))
    thr16 = np.uint64(int(np.float32(p) * np.float32(65536.0)))
    with np.errstate(over="ignore"):
        sseed = site_seed(seed, step, stream)
        idx4 = np.arange((n + 3) // 4, dtype=np.uint64)
        r = _splitmix64(sseed ^ (idx4 * np.uint64(0xD1342543DE82EF95)))
    lanes = np.stack([(r >> np.uint64(16 * k)) & np.uint64(0xFFFF) for k in range(4)], axis=1)
    return (lanes >= thr16).reshape(-1)[:n].reshape(shape)


def keep_mask_from_state(rng_state, stream: int, shape, p: float) -> np.ndarray:
    """``rng_state``: the ``RngState.state`` tensor ({seed, step} as int64)."""
    seed, step = (int(v) for v in rng_state.detach().cpu().tolist())
    return keep_mask(seed, step, stream, shape, p)
