"""The packed weight lanes of filter_score_kernel's two-row loop (csrc/dense4.cuh: pair_word2 / unpack_xy), restated in
numpy: x accumulates the four LUT bytes of docs 4j..4j+3 as one integer (mod 2^32), y the odd bytes in u16 lanes, and the
even sums follow as x - (y << 8).  The identity must hold for every byte sum below 2^16 -- the bound the kernel checks
(`lane_max + exc_sum <= 65535`) before it trusts the lanes."""
import numpy as np
import pytest


def accumulate(rows):
    """rows: [n_rows, 4] uint8 (bytes b0..b3 of the LUT result per row) -> (w0, w1) u32 lanes like unpack_xy"""
    x = np.uint32(0)
    y = np.uint32(0)
    for b in rows.astype(np.uint32):
        p = np.uint32(b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24))
        x = np.uint32((int(x) + int(p)) & 0xFFFFFFFF)                       # x += p
        y = np.uint32((int(y) + int(b[1] | (b[3] << 16))) & 0xFFFFFFFF)     # y += prmt(p, 0, 0x4341)
    e = np.uint32((int(x) - ((int(y) << 8) & 0xFFFFFFFF)) & 0xFFFFFFFF)     # (B2 << 16) | B0
    w0 = np.uint32((int(e) & 0xFFFF) | ((int(y) & 0xFFFF) << 16))           # prmt(e, y, 0x5410)
    w1 = np.uint32((int(e) >> 16) | (int(y) & 0xFFFF0000))                  # prmt(e, y, 0x7632)
    return w0, w1


@pytest.mark.parametrize("n_rows,hi", [(1, 256), (46, 256), (121, 256), (255, 256), (257, 256), (255, 128)])
def test_packed_lanes_match_byte_sums(n_rows, hi):
    rng = np.random.default_rng(n_rows * 7 + hi)
    for _ in range(50):
        rows = rng.integers(0, hi, size=(n_rows, 4), dtype=np.uint8)
        sums = rows.astype(np.uint64).sum(axis=0)
        if sums.max() > 65535:          # the kernel hands such reads to the exact scorer
            continue
        w0, w1 = accumulate(rows)
        assert (int(w0) & 0xFFFF, int(w0) >> 16, int(w1) & 0xFFFF, int(w1) >> 16) == tuple(int(v) for v in sums)


def test_packed_lanes_at_the_limit():
    # 257 rows of 255 in every byte: 65,535 per lane, the largest sum the lanes may hold
    rows = np.full((257, 4), 255, np.uint8)
    w0, w1 = accumulate(rows)
    assert int(w0) == 0xFFFFFFFF and int(w1) == 0xFFFFFFFF
    # one doc saturated, its neighbours empty: no carry leaks into them
    rows = np.zeros((257, 4), np.uint8)
    rows[:, 1] = 255
    w0, w1 = accumulate(rows)
    assert (int(w0) & 0xFFFF, int(w0) >> 16, int(w1)) == (0, 65535, 0)
