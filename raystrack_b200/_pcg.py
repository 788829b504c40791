"""Cranley-Patterson rotation rows without a Python loop.

The reference draws, per emitter and iteration, ``rng = np.random.default_rng(seed + idx_emit + itr)`` followed by
``rng.random(2, dtype=float32)`` and ``rng.random(5, dtype=float32)`` (main.py:1810-1812).  ``main._rotation_table``
needs one such row per distinct sum -- 2041 generator constructions (~18 us each, 40 ms) for the C5 call with a new
seed.  This module evaluates the same arithmetic for all rows at once with NumPy integer arrays:

  * ``SeedSequence(entropy)`` for a single 32-bit entropy word (M. E. O'Neill's ``seed_seq_fe``: hashmix / mix over a
    pool of four uint32, then ``generate_state(4, uint64)``),
  * PCG64 seeding (``pcg_setseq_128_srandom_r``) and four steps of the 128-bit LCG with XSL-RR output,
  * the float32 draw ``(next_uint32 >> 8) * 2**-24``, where ``next_uint32`` hands out the low then the high half of every
    64-bit output.

Bit-identical to NumPy's generators (``tests/test_host_logic.py`` compares 5000 seeds); ``main._rotation_rows`` falls
back to the generator loop for seeds outside [0, 2**32 - rows]."""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_U64 = np.uint64
_INIT_A, _MULT_A = 0x43B0D7E5, 0x931E8875
_INIT_B, _MULT_B = 0x8B51F9DD, 0x58F38DED
_MIX_L, _MIX_R = _U32(0xCA01F9DD), _U32(0x4973F715)
_XS = _U32(16)
_MULT_HI, _MULT_LO = 0x2360ED051FC65DA4, 0x4385DF649FCCF645      # PCG_DEFAULT_MULTIPLIER_128
_M32 = _U64(0xFFFFFFFF)
_S32 = _U64(32)


def _seed_state(entropy: np.ndarray) -> np.ndarray:
    """``SeedSequence(e).generate_state(8)`` (uint32) for every single-word entropy e: array [n, 8]."""
    const = _INIT_A

    def hashmix(v):
        nonlocal const
        v = v ^ _U32(const)
        const = (const * _MULT_A) & 0xFFFFFFFF
        v = v * _U32(const)
        return v ^ (v >> _XS)

    def mix(x, y):
        r = _MIX_L * x - _MIX_R * y
        return r ^ (r >> _XS)

    zero = np.zeros_like(entropy)
    pool = [hashmix(entropy), hashmix(zero), hashmix(zero), hashmix(zero)]
    for src in range(4):
        for dst in range(4):
            if src != dst:
                pool[dst] = mix(pool[dst], hashmix(pool[src]))
    const = _INIT_B
    out = np.empty((entropy.shape[0], 8), _U32)
    for i in range(8):
        v = pool[i & 3] ^ _U32(const)
        const = (const * _MULT_B) & 0xFFFFFFFF
        v = v * _U32(const)
        out[:, i] = v ^ (v >> _XS)
    return out


def _mul64(a: np.ndarray, b: int):
    """Full 128-bit product of uint64 array ``a`` and the 64-bit constant ``b``: (high, low)."""
    a0, a1 = a & _M32, a >> _S32
    b0, b1 = _U64(b & 0xFFFFFFFF), _U64(b >> 32)
    p00, p01, p10, p11 = a0 * b0, a0 * b1, a1 * b0, a1 * b1
    mid = (p00 >> _S32) + (p01 & _M32) + (p10 & _M32)
    return p11 + (p01 >> _S32) + (p10 >> _S32) + (mid >> _S32), (p00 & _M32) | ((mid & _M32) << _S32)


def _step(hi, lo, inc_hi, inc_lo):
    """state = state * MULT + inc  (mod 2**128)."""
    p_hi, p_lo = _mul64(lo, _MULT_LO)
    p_hi = p_hi + lo * _U64(_MULT_HI) + hi * _U64(_MULT_LO)
    new_lo = p_lo + inc_lo
    return p_hi + inc_hi + (new_lo < p_lo).astype(_U64), new_lo


def rotation_rows(seed: int, rows: int) -> np.ndarray:
    """float32 [rows, 7]: row s = the seven draws of ``default_rng(seed + s)`` described above.  ``0 <= seed`` and
    ``seed + rows <= 2**32`` (single-word entropy)."""
    if seed < 0 or seed + rows > (1 << 32):
        raise ValueError("rotation_rows: seed range outside the single-word SeedSequence path")
    with np.errstate(over="ignore"):
        st = _seed_state((np.arange(rows, dtype=np.uint64) + _U64(seed)).astype(_U32)).astype(_U64)
        val = [st[:, 2 * i] | (st[:, 2 * i + 1] << _S32) for i in range(4)]
        init_hi, init_lo = val[0], val[1]
        inc_hi = (val[2] << _U64(1)) | (val[3] >> _U64(63))
        inc_lo = (val[3] << _U64(1)) | _U64(1)
        hi, lo = _step(np.zeros_like(init_hi), np.zeros_like(init_lo), inc_hi, inc_lo)
        lo2 = lo + init_lo
        hi = hi + init_hi + (lo2 < lo).astype(_U64)
        hi, lo = _step(hi, lo2, inc_hi, inc_lo)
        draws = np.empty((rows, 8), _U64)
        for k in range(4):
            hi, lo = _step(hi, lo, inc_hi, inc_lo)
            x = hi ^ lo
            rot = hi >> _U64(58)
            out = (x >> rot) | (x << ((_U64(64) - rot) & _U64(63)))
            draws[:, 2 * k] = out & _M32
            draws[:, 2 * k + 1] = out >> _S32
    return ((draws[:, :7] >> _U64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)
