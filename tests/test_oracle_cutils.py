"""CPU: oracle/cutils_port.c against known answers and (when present) the compiled reference module."""
import numpy as np
import pytest

from oracle import cutils as port
from oracle.ref_loader import reference_available, load_reference_cutils


def test_im2col_event_known_answer_stride1():
    img = np.arange(2 * 4 * 5, dtype=np.float32).reshape(2, 4, 5)
    cols, (oy, ox) = port.im2col_event(img, np.array([0, 3], np.int32), np.array([0, 4], np.int32), 3, 3, 1)
    # (0,0) lies only in window (0,0); (3,4) only in window (1,2)  [Hout=2, Wout=3]
    assert oy.tolist() == [0, 1] and ox.tolist() == [0, 2]
    assert cols.shape == (18, 2) and cols.flags.f_contiguous
    assert np.array_equal(cols[:, 0], img[:, 0:3, 0:3].reshape(-1))
    assert np.array_equal(cols[:, 1], img[:, 1:4, 2:5].reshape(-1))


def test_im2col_event_first_touch_dedup_and_pool_layout():
    img = np.arange(3 * 4 * 4, dtype=np.float32).reshape(3, 4, 4)
    ey = np.array([3, 0, 2, 1], np.int32)
    ex = np.array([3, 0, 2, 1], np.int32)
    cols, (oy, ox) = port.im2col_event(img, ey, ex, 2, 2, 2, chan_as_cols=1)
    assert oy.tolist() == [1, 0] and ox.tolist() == [1, 0]          # windows in first-touch order, deduplicated
    assert cols.shape == (4, 6)
    for site, (wy, wx) in enumerate([(1, 1), (0, 0)]):
        for c in range(3):
            assert np.array_equal(cols[:, site * 3 + c], img[c, 2 * wy:2 * wy + 2, 2 * wx:2 * wx + 2].reshape(-1))


def test_im2col_event_bad_stride():
    with pytest.raises(NotImplementedError):
        port.im2col_event(np.zeros((1, 6, 6), np.float32), np.array([1], np.int32), np.array([1], np.int32), 3, 3, 2)


def test_min_argmax_tie_rules():
    mx = np.asfortranarray(np.array([[1, 5, 2, 2], [1, 5, 7, 2], [0, 5, 7, 2], [1, 1, 7, 2]], np.float32))
    mn = np.asfortranarray(np.array([[3, 2, 0, 4], [2, 2, 5, 4], [0, 1, 4, 3], [1, 0, 4, 4]], np.float32))
    amax, nmin = port.min_argmax(mx, mn)
    # col0: max 1 at rows 0,1,3 -> smaller min_arg wins progressively: row0(3) -> row1(2) -> row3(1); argmin row2 (0) -> unstable
    # col1: max 5 rows 0,1,2: row0(2) -> row1 equal not smaller -> row2(1) wins; argmin row3 (0) -> unstable
    # col2: max 7 rows 1,2,3: row1(5) -> row2(4) -> row3 equal stays row2; argmin row0 (0) -> unstable
    # col3: all equal 2: row0(4) -> row2(3); argmin row2 -> stable
    assert amax.tolist() == [3, 2, 2, 2]
    assert nmin.tolist() == [1, 1, 1, 0]


@pytest.mark.skipif(not reference_available(), reason="/root/reference or oracle/_ref/cutils.so not present")
def test_port_equals_compiled_reference_cutils():
    ref = load_reference_cutils()
    rng = np.random.default_rng(0)
    for trial in range(60):
        c, h, w = int(rng.integers(1, 6)), int(rng.integers(6, 20)) * 2, int(rng.integers(6, 20)) * 2
        img = rng.standard_normal((c, h, w)).astype(np.float32)
        n = int(rng.integers(0, 40))
        ey = rng.integers(0, h, n).astype(np.int32)
        ex = rng.integers(0, w, n).astype(np.int32)
        for (k, s, cac) in [(3, 1, 0), (1, 1, 0), (2, 2, 1), (2, 2, 0), (3, 1, 1), (5, 1, 0)]:
            a_cols, (a_y, a_x) = ref.im2col_event(img, ey, ex, k, k, s, cac)
            b_cols, (b_y, b_x) = port.im2col_event(img, ey, ex, k, k, s, cac)
            assert np.array_equal(a_y, b_y) and np.array_equal(a_x, b_x)
            assert np.array_equal(np.asarray(a_cols), b_cols)
        r, m = int(rng.integers(1, 10)), int(rng.integers(1, 200))
        mx = np.asfortranarray(rng.integers(0, 3, (r, m)).astype(np.float32))
        mn = np.asfortranarray(rng.integers(0, 3, (r, m)).astype(np.float32))
        a, b = ref.min_argmax(mx, mn)
        c2, d = port.min_argmax(mx, mn)
        assert np.array_equal(a, c2) and np.array_equal(b, d)
