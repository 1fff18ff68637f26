"""K1/K2/K3 kernels in isolation: `pa_box_copy` (through the C ABI) against a
NumPy evaluation of the same strided copy -- bit-exact, all element sizes, all
kernel classes (row copy / vector transpose / scalar tile), aligned and odd
shapes, sub-boxes on either side, canaries around the destination."""
import itertools
import math

import numpy as np
import pytest
import torch

from gpu_util import box_copy, np_box_copy, dev_bytes, host_bytes

pytestmark = pytest.mark.gpu

KC_ROWS, KC_TRANSPOSE, KC_TILE = 1, 2, 3


def col_major_strides(dims):
    s, out = 1, []
    for d in dims:
        out.append(s)
        s *= d
    return out


def run_case(extent, sstr, dstr, elsize, n_src, n_dst, src_off=0, dst_off=0, seed=0):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, size=n_src * elsize, dtype=np.uint8)
    dst0 = rng.integers(0, 256, size=n_dst * elsize, dtype=np.uint8)  # canary content
    want = dst0.copy()
    np_box_copy(extent, sstr, dstr, elsize, src, want, src_off, dst_off)
    ds, dd = dev_bytes(src), dev_bytes(dst0)
    desc = box_copy(extent, sstr, dstr, elsize, ds, dd, src_off, dst_off)
    torch.cuda.synchronize()
    got = host_bytes(dd)
    assert got.tobytes() == want.tobytes(), (extent, sstr, dstr, elsize)
    assert host_bytes(ds).tobytes() == src.tobytes()  # source untouched
    return desc


@pytest.mark.parametrize("elsize", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("dims", [(32, 48, 40), (21, 17, 13), (128, 64, 96), (64, 2, 3), (1, 5, 7)])
def test_full_permutations(elsize, dims):
    n = math.prod(dims)
    ss = col_major_strides(dims)
    for perm in itertools.permutations(range(3)):
        # destination has dims (dims[perm[0]], dims[perm[1]], dims[perm[2]]): dst dim i <- src dim perm[i]
        ddims = [dims[p] for p in perm]
        dcol = col_major_strides(ddims)
        ds = [0, 0, 0]
        for i, p in enumerate(perm):
            ds[p] = dcol[i]
        desc = run_case(list(dims), ss, ds, elsize, n, n, seed=sum(perm))
        if perm == (0, 1, 2):
            assert desc.kernel_class == KC_ROWS
        if dims == (128, 64, 96) and perm != (0, 1, 2) and elsize in (4, 8, 16) and perm[0] != 0:
            assert desc.kernel_class == KC_TRANSPOSE and desc.vec_bytes == 16


@pytest.mark.parametrize("elsize", [4, 8, 16])
def test_pack_subbox_to_contiguous(elsize):
    # cfg1-like: box (32,24,32) of (64,24,32) -> contiguous, at two offsets
    parent = (64, 24, 32)
    box = (32, 24, 32)
    for x0 in (0, 32):
        desc = run_case(list(box), col_major_strides(parent), col_major_strides(box), elsize,
                        math.prod(parent), math.prod(box) + 64, src_off=x0, dst_off=32)
        assert desc.kernel_class == KC_ROWS and desc.vec_bytes == 16
    # odd box: falls back to narrower vectors, still exact
    parent, box = (21, 17, 13), (10, 17, 6)
    run_case(list(box), col_major_strides(parent), col_major_strides(box), elsize,
             math.prod(parent), math.prod(box) + 7, src_off=11 + 21 * 17 * 3, dst_off=3)


@pytest.mark.parametrize("ctas", [1, 5, -1])
def test_capped_grid_flavours(ctas):
    """The tile-striding (LOOP=true) flavour of every kernel -- what the NVLink
    put/get paths launch -- must move the same bytes as the one-tile-per-CTA one."""
    from pencilarrays_b200._lib import lib, check
    check(lib.pa_set_tunable(b"box_copy_ctas", ctas))
    try:
        for elsize in (1, 4, 8, 16):
            for dims in ((70, 33, 9), (128, 64, 12)):
                n = math.prod(dims)
                ss = col_major_strides(dims)
                for perm in ((0, 1, 2), (1, 0, 2), (2, 0, 1)):
                    dcol = col_major_strides([dims[p] for p in perm])
                    ds = [0, 0, 0]
                    for i, p in enumerate(perm):
                        ds[p] = dcol[i]
                    run_case(list(dims), ss, ds, elsize, n, n, seed=elsize + sum(perm))
        # narrow row copies (odd sizes -> 8/4/2/1-byte vectors) and a long 1-d run
        run_case([21, 17, 13], col_major_strides((23, 17, 13)), col_major_strides((21, 17, 13)), 8,
                 23 * 17 * 13, 21 * 17 * 13)
        run_case([300001], [1], [1], 2, 300001, 300001)
    finally:
        check(lib.pa_set_tunable(b"box_copy_ctas", 0))


def test_bulk_copy_pipeline_rows():
    """k_rows_bulk (TMA bulk copies through a shared-memory ring) must move the
    same bytes as k_rows: short runs, long runs, partial chunks, few and many CTAs."""
    from pencilarrays_b200._lib import lib, check
    check(lib.pa_set_tunable(b"bulk_rows", 1))
    try:
        for ctas in (0, 1, 7, 148):
            check(lib.pa_set_tunable(b"box_copy_ctas", ctas))
            run_case([1 << 20], [1], [1], 16, 1 << 20, 1 << 20)                   # 16 MiB, 1-d
            run_case([1000], [1], [1], 16, 1000, 1000)                            # 16000 B: < one chunk
            run_case([2049], [1], [1], 16, 2049, 2049)                            # 2 chunks + 16 B
            parent, box = (64, 24, 32), (32, 24, 32)                              # 512-B runs
            run_case(list(box), col_major_strides(parent), col_major_strides(box), 16,
                     math.prod(parent), math.prod(box) + 8, src_off=32, dst_off=8)
            parent, box = (4096, 6, 5), (2050, 6, 3)                              # 32800-B runs: 3 chunks/row
            desc = run_case(list(box), col_major_strides(parent), col_major_strides(parent), 16,
                            math.prod(parent), math.prod(parent), src_off=5, dst_off=4096 * 6 + 7)
            assert desc.kernel_class == KC_ROWS
            run_case([256, 37], [1, 300], [1, 256], 4, 300 * 37, 256 * 37)        # f32, 1 KiB runs
    finally:
        check(lib.pa_set_tunable(b"bulk_rows", 0))
        check(lib.pa_set_tunable(b"box_copy_ctas", 0))


@pytest.mark.parametrize("elsize", [4, 8, 16])
@pytest.mark.parametrize("perm", [(1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0), (0, 2, 1)])
def test_unpack_contiguous_to_permuted_subbox(elsize, perm):
    src_dims = (32, 24, 40)
    box = [src_dims[p] for p in perm]            # dst box dims
    parent = [box[0] + 32, box[1], box[2] + 8]   # dst parent is larger than the box
    pcol = col_major_strides(parent)
    ds = [0, 0, 0]
    for i, p in enumerate(perm):
        ds[p] = pcol[i]
    off = 16 + pcol[2] * 3
    desc = run_case(list(src_dims), col_major_strides(src_dims), ds, elsize,
                    math.prod(src_dims), math.prod(parent), dst_off=off)
    if perm[0] != 0:
        assert desc.kernel_class == KC_TRANSPOSE


@pytest.mark.parametrize("elsize", [1, 2, 4, 8, 16])
def test_degenerate_shapes(elsize):
    run_case([1], [1], [1], elsize, 4, 4, src_off=1, dst_off=2)            # one element
    run_case([0, 5], [1, 3], [1, 7], elsize, 16, 40)                        # empty box
    run_case([1000003], [1], [1], elsize, 1000003, 1000003)                 # long odd 1-d run
    run_case([7, 1, 9], [1, 7, 7], [9, 1, 1], elsize, 63, 63)               # extent-1 dims
    run_case([5, 3], [2, 10], [1, 5], elsize, 30, 15)                       # non-unit source stride
    run_case([5, 3], [1, 5], [3, 1], elsize, 15, 15)                        # tiny transpose
    run_case([4, 3], [1, 4], [2, 8], elsize, 12, 24)                        # non-unit dest stride


@pytest.mark.parametrize("elsize", [4, 16])
def test_extra_dims_5d(elsize):
    # (a,b,c) spatial + extras (3,4): spatial permuted (2,0,1), extras identity
    dims = (16, 12, 10, 3, 4)
    ss = col_major_strides(dims)
    ddims = (dims[2], dims[0], dims[1], 3, 4)
    dcol = col_major_strides(ddims)
    ds = [dcol[1], dcol[2], dcol[0], dcol[3], dcol[4]]
    run_case(list(dims), ss, ds, elsize, math.prod(dims), math.prod(dims))


def test_partial_tiles_every_residue():
    # extents that leave every kind of partial tile in the vector transpose (S=8: 64x64 tiles)
    for ex, ey in [(64, 64), (66, 64), (64, 70), (2, 2), (126, 2), (2, 130), (190, 66)]:
        dims = (ex, 3, ey)
        ss = col_major_strides(dims)
        dcol = col_major_strides((ey, 3, ex))
        ds = [dcol[2], dcol[1], dcol[0]]
        desc = run_case(list(dims), ss, ds, 8, math.prod(dims), math.prod(dims), seed=ex + ey)
        assert desc.kernel_class == KC_TRANSPOSE and desc.vec_bytes == 16
