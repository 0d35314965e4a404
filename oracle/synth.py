"""Counter-based synthetic embedding generator -- numpy restatement.

TEST INFRASTRUCTURE (see oracle/vm_oracle.c header).  Same definition as
`vo_synth_rows` (oracle/vm_oracle.c) and the device generator (csrc/synth.cuh):
value(seed,row,col) = (byte(col&7 of splitmix64(rowkey + col>>3)) * 255 >> 8) - 127) / 128,
rowkey = splitmix64(seed * GOLDEN + row).  All values are multiples of 1/128 in
[-127/128, 127/128]: exactly representable in bf16, tf32 and fp32, and every 384/768-term dot
product of them is exact in fp32, so quantisation cannot perturb parity.
"""
from __future__ import annotations

import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_DUPK = np.uint64(0xD6E8FEB86659FD93)
_DUPC = np.uint64(0x632BE59BD9B4E019)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (np.asarray(x, dtype=np.uint64) + _G).astype(np.uint64)
        z = x
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _base_rows(seed: int, rows: np.ndarray, d: int) -> np.ndarray:
    rows = np.asarray(rows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        rowkey = splitmix64(np.uint64(seed) * _G + rows)  # [n]
        words = splitmix64(rowkey[:, None] + np.arange((d + 7) // 8, dtype=np.uint64)[None, :])
    b = words.view(np.uint8).reshape(len(rows), -1)[:, :d].astype(np.int32)  # little-endian bytes
    return (((b * 255) >> 8) - 127).astype(np.float32) / np.float32(128.0)


def synth_rows(seed: int, row0: int, n: int, d: int, dup_period: int = 0) -> np.ndarray:
    """Rows [row0, row0+n) of the synthetic store as float32 [n, d]."""
    rows = np.arange(row0, row0 + n, dtype=np.uint64)
    return synth_rows_at(seed, rows, d, dup_period)


def synth_rows_at(seed: int, rows: np.ndarray, d: int, dup_period: int = 0) -> np.ndarray:
    rows = np.asarray(rows, dtype=np.uint64)
    out = _base_rows(seed, rows, d)
    if dup_period > 0 and len(rows):
        with np.errstate(over="ignore"):
            hk = splitmix64(np.array([np.uint64(seed) ^ _DUPK], dtype=np.uint64))[0]
            hr = splitmix64(hk + rows)
            planted = (rows > 0) & (hr % np.uint64(dup_period) == 0)
            if planted.any():
                pr = rows[planted]
                hrp = hr[planted]
                parent = splitmix64(hrp) % pr
                words = splitmix64((hrp + _DUPC)[:, None] + np.arange((d + 7) // 8, dtype=np.uint64)[None, :])
                b = words.view(np.uint8).reshape(len(pr), -1)[:, :d]
                copy = (b & 15) != 0
                pv = _base_rows(seed, parent, d)
                sub = out[planted]
                sub[copy] = pv[copy]
                out[planted] = sub
    return out


def synth_queries(seed: int, q: int, d: int, store_seed: int, n_store: int, dup_period: int = 0) -> np.ndarray:
    """Query batch: even queries are perturbed copies of a store row (1/4 of the columns
    redrawn) so the top hit is non-trivial; odd queries are independent draws."""
    out = _base_rows(seed, np.arange(q, dtype=np.uint64), d)
    if n_store > 0:
        with np.errstate(over="ignore"):
            h = splitmix64(np.uint64(seed) * _G + np.arange(q, dtype=np.uint64) + np.uint64(0x1234567))
        target = h % np.uint64(n_store)
        even = (np.arange(q) % 2) == 0
        src = synth_rows_at(store_seed, target[even], d, dup_period)
        keep = (np.arange(d) % 4) != 0
        sub = out[even]
        sub[:, keep] = src[:, keep]
        out[even] = sub
    return out
