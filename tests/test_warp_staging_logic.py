"""Host-side model of the staged warp's integer logic (csrc/warp.cu: warp_fwd_staged_kernel) checked by property tests on CPU:

* the persistent tile walk -- `advance` (carries, no divisions) must visit exactly the tiles `first(t + G)` names, for every
  tiling / grid size, including grids larger than one image and single-column tilings;
* the window plan -- whenever the kernel decides that a tile is served from a staging window, every corner of every valid pixel
  lies inside that window (rows and columns), and the window's first column is 16-byte aligned for 2- and 4-byte frames.

The restatements below follow the kernel line by line; the GPU suite checks the kernel's VALUES against the L1 kernel, this
file checks the index arithmetic on far more shapes than a GPU test can afford."""
import math

import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

WS_TILE, WS_BW, WS_SMALL_W, WS_SMALL_H = 32, 64, 48, 40


def first(t, tiles_x, tiles_y):
    per_img = tiles_x * tiles_y
    b = t // per_img
    r = t - b * per_img
    ty = r // tiles_x
    return b, ty, r - ty * tiles_x


def advance(o, G, tiles_x, tiles_y):
    per_img = tiles_x * tiles_y
    gb = G // per_img
    gy = (G - gb * per_img) // tiles_x
    gx = G - gb * per_img - gy * tiles_x
    b, ty, tx = o
    tx += gx
    if tx >= tiles_x:
        tx -= tiles_x
        ty += 1
    ty += gy
    if ty >= tiles_y:
        ty -= tiles_y
        b += 1
    b += gb
    return b, ty, tx


@settings(max_examples=400, deadline=None)
@given(st.integers(1, 70), st.integers(1, 40), st.integers(1, 5), st.integers(1, 700), st.integers(0, 5))
def test_tile_walk_by_carries_equals_division(tiles_x, tiles_y, B, G, steps):
    N = tiles_x * tiles_y * B
    for t0 in {0, min(G, N) - 1, (G // 2) % N}:
        o, t = first(t0, tiles_x, tiles_y), t0
        for _ in range(steps):
            o, t = advance(o, G, tiles_x, tiles_y), t + G
            if t < N:                                     # beyond the last tile the origin is computed but never used
                assert o == first(t, tiles_x, tiles_y), (tiles_x, tiles_y, G, t)
                assert 0 <= o[1] < tiles_y and 0 <= o[2] < tiles_x and o[0] < B


def plan(x0, y0, valid, big_h):
    """plan_tile: bounding box of the valid pixels' north-west corners -> (anchor x, anchor y, window w, window h); w = 0: L1 path."""
    mnx, mxx, mny, mxy = x0[valid].min(), x0[valid].max(), y0[valid].min(), y0[valid].max()
    mnx = (int(mnx) >> 3) << 3                            # floor to a multiple of 8 (arithmetic shift: also for negatives)
    ex, ey = int(mxx) - mnx + 2, int(mxy) - int(mny) + 2
    if ex <= WS_SMALL_W and ey <= WS_SMALL_H:
        return mnx, int(mny), WS_SMALL_W, WS_SMALL_H
    if ex <= WS_BW and ey <= big_h:
        return mnx, int(mny), WS_BW, big_h
    return mnx, int(mny), 0, 0


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([0.03, 0.6, 3.0, 9.0, 30.0]), st.floats(-0.4, 0.4), st.floats(-0.4, 0.4),
       st.floats(-300.0, 300.0), st.floats(-300.0, 300.0), st.sampled_from([(2, 64), (4, 32)]), st.integers(1, 32), st.integers(1, 16))
def test_a_staged_tile_holds_every_corner_it_gathers(seed, sigma, shear_x, shear_y, tx, ty, fmt, rows, col_pairs):
    es, big_h = fmt
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.arange(WS_TILE), np.arange(WS_TILE), indexing="ij")
    valid = (ys < rows) & (xs < 2 * col_pairs)            # tiles at the frame edge: whole pixel pairs, whole rows
    px = 1000 + xs + tx + shear_x * (xs + ys) + sigma * rng.standard_normal((WS_TILE, WS_TILE))
    py = 1000 + ys + ty + shear_y * (xs - ys) + sigma * rng.standard_normal((WS_TILE, WS_TILE))
    x0, y0 = np.floor(px).astype(np.int64) - 1000, np.floor(py).astype(np.int64) - 1000      # corners may be negative
    ax, ay, bw, bh = plan(x0, y0, valid, big_h)
    assert (ax * es) % 16 == 0                            # the copy engine's alignment rule for the innermost coordinate
    assert ax <= x0[valid].min() < ax + 8
    if bw:
        o = (y0 - ay) * bw + (x0 - ax)                    # the kernel's gather offset of the north-west corner
        ov = o[valid]
        assert ov.min() >= 0 and (ov + bw + 1).max() < bw * bh       # ... and of the south-east one, inside one plane
        cx = (x0 - ax)[valid]
        assert cx.min() >= 0 and cx.max() + 1 <= bw - 1              # no corner wraps into the next window row
    else:                                                 # refused only when the box really does not fit the big window
        assert x0[valid].max() - ax + 2 > WS_BW or y0[valid].max() - y0[valid].min() + 2 > big_h


def test_near_identity_flow_always_takes_the_small_window():
    """The flow the reference's model produces (|f| << 1 px): a 32 x 32 tile's corners span 33-34 columns from an anchor at most 8
    pixels to the left -- 48 x 40 always suffices, wherever the tile sits."""
    rng = np.random.default_rng(0)
    for tile_x in range(0, 1920, 32):
        ys, xs = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
        f = 0.03 * rng.standard_normal((2, 32, 32))
        x0 = np.floor(tile_x + xs + f[0]).astype(np.int64)
        y0 = np.floor(64 + ys + f[1]).astype(np.int64)
        _, _, bw, bh = plan(x0, y0, np.ones((32, 32), bool), 64)
        assert (bw, bh) == (WS_SMALL_W, WS_SMALL_H)
    assert math.ceil(1920 / 32) * math.ceil(1080 / 32) * 8 == 16320          # the tile count of cfg2 quoted in DESIGN.md
