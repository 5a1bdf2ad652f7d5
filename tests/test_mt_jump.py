"""Host side of the MT19937 jump-ahead (speechsplit_b200/csrc/mt_jump.cpp), checked on the CPU against
numpy's own generator: the characteristic polynomial, x^J mod phi, and the tap-list identity
w[n + J] = XOR_{taps i} w[n + i] that the GPU kernel evaluates."""
import numpy as np
from numpy.random import RandomState

from speechsplit_b200 import _lib


def _lib_handle():
    return _lib.load()


def _untempered(seed, n_blocks):
    """w[0 .. 624 (n_blocks + 1)): the seeded state followed by n_blocks regenerated state blocks."""
    rs = RandomState(seed)
    out = [rs.get_state()[1].copy()]
    for _ in range(n_blocks):
        rs.randint(0, 2 ** 32, size=624, dtype=np.uint32)        # consumes exactly one block
        st = rs.get_state()
        assert st[2] == 624
        out.append(st[1].copy())
    return np.concatenate(out)


def _state_after(seed, n_blocks):
    rs = RandomState(seed)
    left = n_blocks
    while left:
        step = min(left, 4096)
        rs.randint(0, 2 ** 32, size=624 * step, dtype=np.uint32)
        left -= step
    return rs.get_state()[1]


def _apply(seed, taps):
    w = _untempered(seed, 34)
    acc = np.zeros(624, np.uint32)
    for i in taps:
        acc ^= w[i:i + 624]
    return acc


def _same_state(a, b):
    # the low 31 bits of the oldest word are not part of the 19937-bit state
    return np.array_equal(a[1:], b[1:]) and (a[0] >> 31) == (b[0] >> 31)


def test_characteristic_polynomial():
    lib = _lib_handle()
    n = lib.ssfe_mt_charpoly_terms(None, 0)
    assert n == 134                      # 135 terms with x^19937 (Matsumoto & Nishimura 1998, table II)
    exps = np.zeros(n, np.int32)
    lib.ssfe_mt_charpoly_terms(exps.ctypes.data, n)
    assert exps.min() == 0 and exps.max() < 19937 - 64 and len(set(exps.tolist())) == n


def test_small_powers_are_monomials():
    lib = _lib_handle()
    poly = np.zeros(312, np.uint64)
    for e in (0, 1, 63, 64, 19936):
        assert lib.ssfe_mt_jump_poly(e, poly.ctypes.data) == 0
        bits = np.unpackbits(poly.view(np.uint8), bitorder="little")
        assert bits.sum() == 1 and bits[e] == 1


def test_jump_taps_reproduce_numpy_states():
    lib = _lib_handle()
    unit = 624 * 4096
    taps = np.zeros(20000, np.uint16)
    for level, d, seed in ((0, 1, 226), (0, 2, 7), (0, 9, 4000000000)):
        n = lib.ssfe_mt_jump_taps(unit, level, d, taps.ctypes.data, taps.size)
        assert 9000 < n < 11000
        assert np.all(np.diff(taps[:n].astype(np.int64)) > 0)
        assert _same_state(_apply(seed, taps[:n].astype(np.int64)), _state_after(seed, 4096 * d * 256 ** level))


def test_jump_taps_second_level():
    lib = _lib_handle()
    taps = np.zeros(20000, np.uint16)
    n = lib.ssfe_mt_jump_taps(624 * 16, 1, 3, taps.ctypes.data, taps.size)   # x^(3 * 256 * 16 blocks)
    assert _same_state(_apply(31, taps[:n].astype(np.int64)), _state_after(31, 3 * 256 * 16))
    # the unit is part of the cache key: asking for another unit rebuilds the table
    n = lib.ssfe_mt_jump_taps(624 * 4096, 0, 1, taps.ctypes.data, taps.size)
    assert _same_state(_apply(31, taps[:n].astype(np.int64)), _state_after(31, 4096))
