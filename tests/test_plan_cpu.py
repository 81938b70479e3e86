"""CPU: host-side launch planning of the tcgen05 kernels (`ustrun_tc_plan_query`, csrc/tc_conv.cu) -- the tile
planner for small feature maps and the split / rows-per-stage choice of the row-mode weight gradient.  No kernel is
launched; without a GPU the planner assumes the 148 SMs of a B200."""
import ctypes
import itertools
import os

import pytest

from conftest import PKG

SMS = 148
MAX_PARTS = 1280


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(os.path.join(PKG, "libustrun_sm100.so"))


def _q(lib, what, B, H, W, cin, cout, ks=3):
    out = (ctypes.c_int * 8)()
    assert lib.ustrun_tc_plan_query(what, B, H, W, cin, cout, ks, out) == 0
    return list(out)


def _cdiv(a, b):
    return -(-a // b)


def test_bottleneck_level_of_cfg2_uses_two_box_tiles(lib):
    # 24x24 feature maps of a 384x384 input: 8x8 boxes cover them exactly -> 36 full tiles = one round
    for B, cout, want_bn, want_tiles in ((8, 1024, 256, 36), (8, 512, 128, 36), (16, 512, 256, 72), (24, 1024, 256, 108)):
        bn, row, boxes, tw, th, m_tiles, per, n_tiles = _q(lib, 0, B, 24, 24, 1024, cout)
        assert (bn, row, boxes, tw, th, m_tiles) == (want_bn, 0, 2, 8, 8, want_tiles)
        assert n_tiles * per <= SMS
    # large feature maps keep one box per tile and the wide N tile; the 384-wide 64/128-channel layers run in row mode
    assert _q(lib, 0, 8, 96, 96, 256, 256)[:5] == [256, 0, 1, 32, 4]
    assert _q(lib, 0, 8, 48, 48, 512, 512)[:3] == [256, 0, 1]
    assert _q(lib, 0, 8, 384, 384, 64, 64)[:5] == [64, 1, 1, 128, 1]
    assert _q(lib, 0, 8, 384, 384, 64, 64, ks=1)[1] == 0            # 1x1: never row mode


def test_forward_plan_invariants(lib):
    for B, H, W, cout in itertools.product((1, 3, 8, 16), (5, 12, 18, 24, 40), (8, 16, 24, 30, 64, 192), (64, 128, 320, 512, 2048)):
        _check_forward_plan(lib, B, H, W, cout)


def _check_forward_plan(lib, B, H, W, cout):
    bn, row, boxes, tw, th, m_tiles, per, n_tiles = _q(lib, 0, B, H, W, 64, cout)
    assert cout % bn == 0 and n_tiles == cout // bn and bn in (64, 128, 256)
    assert boxes in (1, 2) and 1 <= tw <= W and 1 <= th <= H and tw * th * boxes <= 128
    units = B * _cdiv(W, tw) * _cdiv(H, th)
    assert m_tiles == _cdiv(units, boxes)                             # every pixel belongs to exactly one box
    assert per == max(1, min(SMS // n_tiles if n_tiles <= SMS else 1, m_tiles, MAX_PARTS))
    if bn < (256 if cout % 256 == 0 else (128 if cout % 128 == 0 else 64)):
        assert bn == 128                                              # the only narrowing the planner does: 256 -> 128


def test_row_wgrad_plan_invariants(lib):
    for B, H, W in itertools.product((1, 8, 24), (6, 24, 48), (16, 18, 24, 32, 48, 64, 96, 192, 384)):
        for cin, cout in ((64, 64), (128, 64), (512, 1024), (1024, 1024), (1024, 512)):
            _check_row_wgrad_plan(lib, B, H, W, cin, cout)


def _check_row_wgrad_plan(lib, B, H, W, cin, cout):
    ok, swap, bn, tw, R, nsegs, splits, items = _q(lib, 1, B, H, W, cin, cout)
    assert ok == 1
    assert tw % 16 == 0 and tw <= 64 and _cdiv(W, tw) * tw >= W
    assert R >= 1 and (R == 1 or R * tw <= 64) and R <= H
    assert nsegs == B * _cdiv(H, R) * _cdiv(W, tw)
    assert 1 <= splits <= nsegs
    assert swap == (1 if cout == 64 and cin >= 128 else 0)
    m_side, n_side = (cin, cout) if swap else (cout, cin)
    assert bn == (128 if n_side % 128 == 0 else 64) and items == 3 * _cdiv(m_side, 128) * (n_side // bn)


def test_row_wgrad_waves_and_rows_per_stage_at_cfg2(lib):
    # 96 / 192 work items on 148 SMs: extra K splits even the waves out when the reduce cost allows it
    assert _q(lib, 1, 8, 48, 48, 1024, 512)[6] == 3 and _q(lib, 1, 16, 24, 24, 512, 1024)[6] == 3
    assert _q(lib, 1, 8, 24, 24, 1024, 1024)[6] == 2 and _q(lib, 1, 8, 24, 24, 512, 1024)[6] == 1
    # narrow feature maps: two 32-pixel (or four 16-pixel) image rows per stage
    assert _q(lib, 1, 8, 24, 24, 1024, 1024)[3:5] == [32, 2] and _q(lib, 1, 8, 16, 16, 1024, 1024)[3:5] == [16, 4]
    assert _q(lib, 1, 8, 384, 384, 64, 64)[3:7] == [64, 1, 8 * 384 * 6, 49]


def test_plan_query_rejects_bad_arguments(lib):
    out = (ctypes.c_int * 8)()
    assert lib.ustrun_tc_plan_query(0, 8, 24, 24, 60, 64, 3, out) != 0          # channels must be multiples of 64
    assert lib.ustrun_tc_plan_query(2, 8, 24, 24, 64, 64, 3, out) != 0
    assert lib.ustrun_tc_plan_query(0, 8, 24, 24, 64, 64, 3, None) != 0
