"""Round-2 additions, through the C ABI, bit-exact:

* element-aligned transposes (odd extents -- the N/2+1 grids of real-to-complex
  transforms): the vector transpose kernel with element-wise source and/or
  destination side instead of the scalar tile;
* the one-launch multi-peer put / get kernel (`pa_put_all` / `pa_get_all`) with
  the peers emulated by local arrays, against the oracle;
* the host paths: `pa_transpose_host` (cut + pipelined upload/kernel/download) and
  `pa_host_chain_*` (asynchronous, double-buffered), against the device path and
  the oracle;
* programmatic dependent launch on/off gives identical bytes.
"""
import ctypes as C
import itertools
import math

import numpy as np
import pytest
import torch

import pencilarrays_b200 as pa
from pencilarrays_b200._lib import lib, check
from pencilarrays_b200.transpositions import _Plan
from oracle import pencil_oracle as O
from util import CASES, DTYPES, build_chain
from gpu_util import dev_bytes, host_bytes, ptr, stream_ptr
from test_gpu_kernels import run_case, col_major_strides, KC_TRANSPOSE, KC_ROWS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("elsize", [4, 8])
@pytest.mark.parametrize("dims", [(65, 32, 12), (64, 33, 12), (65, 33, 12), (513, 64, 3),
                                  (130, 3, 67), (1, 65, 33)])
def test_element_aligned_transposes(elsize, dims):
    """Odd X extent -> element-wise loads (SE), odd Y extent -> element-wise stores
    (DE); the tile, the shared-memory layout and the other side stay 128-bit."""
    n = math.prod(dims)
    ss = col_major_strides(dims)
    for perm in ((1, 0, 2), (2, 0, 1), (1, 2, 0), (2, 1, 0)):
        ddims = [dims[p] for p in perm]
        dcol = col_major_strides(ddims)
        ds = [0, 0, 0]
        for i, p in enumerate(perm):
            ds[p] = dcol[i]
        desc = run_case(list(dims), ss, ds, elsize, n, n, seed=sum(perm) + elsize)
        if dims[0] > 1 and dims[perm[0]] > 1:
            assert desc.kernel_class == KC_TRANSPOSE
            if perm == (1, 0, 2):  # nothing merges: X = dims[0], Y = dims[1]
                odd = (dims[0] * elsize) % 16 != 0 or (dims[1] * elsize) % 16 != 0
                assert desc.vec_bytes == (elsize if odd else 16)


@pytest.mark.parametrize("elsize", [4, 8])
def test_element_aligned_subboxes_and_offsets(elsize):
    # source rows start at odd element offsets inside a larger parent; destination likewise
    parent_s, box = (71, 40, 9), (64, 32, 8)
    parent_d = (35, 70, 9)  # dst dims (y, x, z) with room around the box
    dcol = col_major_strides(parent_d)
    for so, do in ((0, 0), (3, 0), (0, 5), (3 + 71 * 2, 1 + 35 * 3)):
        desc = run_case(list(box), col_major_strides(parent_s), [dcol[1], dcol[0], dcol[2]], elsize,
                        math.prod(parent_s), math.prod(parent_d), src_off=so, dst_off=do)
        assert desc.kernel_class == KC_TRANSPOSE
    # narrow row copies now keep 16 accesses in flight per thread: odd run lengths
    for ex in (513, 1025, 21):
        parent = (ex + 6, 9, 5)
        desc = run_case([ex, 9, 5], col_major_strides(parent), col_major_strides((ex, 9, 5)), elsize,
                        math.prod(parent), ex * 45, src_off=3)
        assert desc.kernel_class == KC_ROWS


def test_pdl_on_off_same_bytes():
    for v in (0, 1):
        check(lib.pa_set_tunable(b"pdl", v))
        try:
            for _ in range(3):  # back-to-back launches on one stream
                run_case([128, 64, 12], col_major_strides((128, 64, 12)),
                         [64 * 12, 1, 64], 8, 128 * 64 * 12, 128 * 64 * 12)
        finally:
            check(lib.pa_set_tunable(b"pdl", 1))


@pytest.mark.parametrize("case", [c for c in CASES if math.prod(c["grid"]) > 1],
                         ids=[c["name"] for c in CASES if math.prod(c["grid"]) > 1])
@pytest.mark.parametrize("cap", [0, 3])
def test_multi_peer_put_get_one_launch(case, cap):
    """`pa_put_all` / `pa_get_all`: every remote block of a rank in ONE launch
    (tiles interleaved over the peers), the peers' arrays emulated on this GPU."""
    dtype, it, extra = DTYPES[case["it"]], case["it"], case["extra"]
    ranks, steps = build_chain(case)
    g = O.global_pattern(case["dims"], extra, it)
    cur_o = O.scatter(g, [po for (_, po) in steps[0]], extra, dtype)
    st = stream_ptr()
    for k in range(1, len(steps)):
        nxt_o = [O.OArray.undef(dtype, po, *extra) for (_, po) in steps[k]]
        O.transpose_all(nxt_o, cur_o)
        plans = [_Plan(steps[k - 1][r][0], steps[k][r][0], extra, it, pa.PeerPut())
                 for r in range(len(ranks))]
        if plans[0].info.dim != 0:
            cur = [dev_bytes(a.data.reshape(-1, order="F")) for a in cur_o]
            for flavour in ("put", "get"):
                nxt = [torch.full((max(1, a.data.size * it),), 0xA5, dtype=torch.uint8, device="cuda")
                       for a in nxt_o]
                n0 = pa.launch_count()
                for r, pl in enumerate(plans):
                    nproc = pl.info.nproc
                    check(lib.pa_copy_self(pl.h, ptr(cur[r]), ptr(nxt[r]), st))
                    arr = (C.c_void_p * nproc)()
                    for p in range(1, nproc + 1):
                        peer = pl.peer(p)
                        arr[p - 1] = (nxt if flavour == "put" else cur)[peer.world_rank].data_ptr()
                    if flavour == "put":
                        check(lib.pa_put_all(pl.h, ptr(cur[r]), arr, cap, st))
                    else:
                        check(lib.pa_get_all(pl.h, arr, ptr(nxt[r]), cap, st))
                torch.cuda.synchronize()
                for r, a in enumerate(nxt_o):
                    want = np.ascontiguousarray(a.data.reshape(-1, order="F")).view(np.uint8)
                    assert host_bytes(nxt[r])[:want.size].tobytes() == want.tobytes(), (flavour, k, r)
        cur_o = nxt_o


# ---------------------------------------------------------------------------- host paths
def _single_rank_chain(dims, dtype, perms):
    topo = pa.MPITopology(pa.COMM_SELF, (1, 1))
    px = pa.Pencil(topo, dims, (2, 3))
    py = pa.Pencil(px, decomp_dims=(1, 3), permute=pa.Permutation(*perms[0]))
    pz = pa.Pencil(py, decomp_dims=(1, 2), permute=pa.Permutation(*perms[1]))
    ux, uy, uz = (pa.PencilArray.undef(dtype, p) for p in (px, py, pz))
    return (px, py, pz), (ux, uy, uz)


@pytest.mark.parametrize("chunk", [1 << 12, 1 << 16, 1 << 40])
@pytest.mark.parametrize("perms", [((2, 1, 3), (3, 2, 1)), ((2, 3, 1), (3, 1, 2))])
def test_transpose_host_equals_device_path(chunk, perms):
    """`pa_transpose_host`: cut along the outermost source dim, upload || kernel ||
    download -- same bytes as the device-resident transpose!, whatever the cut."""
    check(lib.pa_set_tunable(b"host_chunk_bytes", chunk))
    try:
        dims = (40, 24, 36)
        _, (ux, uy, uz) = _single_rank_chain(dims, torch.float64, perms)
        ux.data.normal_()
        for (dst, src) in ((uy, ux), (uz, uy), (uy, uz), (ux, uy)):
            t = pa.Transposition(dst, src)
            pa.transpose_(t)
            torch.cuda.synchronize()
            hin = src.data.cpu().pin_memory()
            hout = torch.empty_like(hin).reshape(dst.data.shape).pin_memory()
            hout.view(torch.uint8).fill_(0x77)
            pa.transpose_host_(t, hin, hout)
            assert hout.view(torch.uint8).numpy().tobytes() == \
                dst.data.cpu().view(torch.uint8).numpy().tobytes()
    finally:
        check(lib.pa_set_tunable(b"host_chunk_bytes", 64 << 20))


@pytest.mark.parametrize("nplans", [1, 2, 4])
@pytest.mark.parametrize("chunk", [1 << 13, 1 << 40])
def test_host_chain_async_double_buffered(nplans, chunk):
    """`pa_host_chain_*`: more submits in flight than device slots, distinct inputs,
    every result equal to the device-resident chain."""
    check(lib.pa_set_tunable(b"host_chunk_bytes", chunk))
    try:
        dims = (32, 20, 28)
        _, (ux, uy, uz) = _single_rank_chain(dims, torch.complex128, ((2, 1, 3), (3, 2, 1)))
        pairs = [(uy, ux), (uz, uy), (uy, uz), (ux, uy)][:nplans]
        ts = [pa.Transposition(d, s) for d, s in pairs]
        chain = pa.HostChain(ts)
        last = pairs[-1][0]
        ins, outs, wants = [], [], []
        for i in range(5):
            ux.data.view(torch.float64).normal_()
            ins.append(ux.data.cpu().pin_memory())
            for t in ts:
                pa.transpose_(t)
            torch.cuda.synchronize()
            wants.append(last.data.cpu().view(torch.uint8).numpy().tobytes())
            outs.append(torch.empty(last.data.shape, dtype=torch.complex128).pin_memory())
        tickets = [chain.submit(a, b) for a, b in zip(ins, outs)]
        assert tickets == list(range(5))
        chain.wait(tickets[1])
        assert outs[0].view(torch.uint8).numpy().tobytes() == wants[0]
        assert outs[1].view(torch.uint8).numpy().tobytes() == wants[1]
        chain.wait()
        for o, w in zip(outs, wants):
            assert o.view(torch.uint8).numpy().tobytes() == w
    finally:
        check(lib.pa_set_tunable(b"host_chunk_bytes", 64 << 20))


def test_host_paths_against_oracle():
    """Host in, host out, no device array of ours in between: oracle bytes."""
    dims = (16, 21, 41)
    topo = pa.MPITopology(pa.COMM_SELF, (1, 1))
    px = pa.Pencil(topo, dims, (2, 3))
    py = pa.Pencil(px, decomp_dims=(1, 3), permute=pa.Permutation(2, 3, 1))
    ox = O.OPencil(O.OTopology((1, 1), 0), dims, (2, 3))
    oy = O.OPencil(O.OTopology((1, 1), 0), dims, (1, 3), (2, 3, 1))
    g = O.global_pattern(dims, (), 8)
    (ax,) = O.scatter(g, [ox], (), np.float64)
    ay = O.OArray.undef(np.float64, oy)
    O.transpose_all([ay], [ax])
    ux, uy = pa.PencilArray.undef(torch.float64, px), pa.PencilArray.undef(torch.float64, py)
    t = pa.Transposition(uy, ux)
    hin = torch.from_numpy(np.ascontiguousarray(ax.data.reshape(-1, order="F")).copy())
    hout = torch.empty(ay.data.size, dtype=torch.float64)
    check(lib.pa_set_tunable(b"host_chunk_bytes", 4096))
    try:
        pa.transpose_host_(t, hin, hout)  # pageable host memory works too
        want = np.ascontiguousarray(ay.data.reshape(-1, order="F")).view(np.uint8).tobytes()
        assert hout.view(torch.uint8).numpy().tobytes() == want
        chain = pa.HostChain([t])
        hout.zero_()
        chain.wait(chain.submit(hin, hout))
        assert hout.view(torch.uint8).numpy().tobytes() == want
    finally:
        check(lib.pa_set_tunable(b"host_chunk_bytes", 64 << 20))


def test_empty_local_array_null_pointers():
    """A rank that owns nothing passes NULL: accepted (reference: more processes than
    points, Pencils.jl:193-218); a NULL for a non-empty array is still an error."""
    topo = pa.MPITopology(pa.COMM_SELF, (1, 1))
    px = pa.Pencil(topo, (0, 4, 4), (2, 3))
    py = pa.Pencil(px, decomp_dims=(1, 3), permute=pa.Permutation(2, 1, 3))
    ux, uy = pa.PencilArray.undef(torch.float32, px), pa.PencilArray.undef(torch.float32, py)
    pa.transpose_(uy, ux)  # nothing to move, no error
    qx = pa.Pencil(topo, (4, 4, 4), (2, 3))
    qy = pa.Pencil(qx, decomp_dims=(1, 3), permute=pa.Permutation(2, 1, 3))
    plan = _Plan(qx, qy, (), 4, pa.PointToPoint())
    st = lib.pa_transpose(plan.h, None, None, None, 1, stream_ptr())
    assert st == pa._lib.PA_EINVAL


# ---------------------------------------------------------------------------- fused unpack + FFT
def test_fft_core_on_cpu_matches_numpy(tmp_path):
    """(runs the shared butterfly / indexing core on the host -- see also the CPU suite)"""
    import test_fft_core
    test_fft_core.test_fft_core_matches_numpy(tmp_path)


@pytest.mark.parametrize("dims,perms", [((8, 64, 4), ((2, 1, 3), (3, 2, 1))),
                                        ((24, 8, 16), ((2, 1, 3), (3, 2, 1))),
                                        ((12, 256, 8), ((2, 3, 1), (3, 1, 2))),
                                        ((9, 512, 3), ((2, 1, 3), (3, 2, 1))),
                                        ((16, 1024, 8), ((2, 3, 1), (3, 1, 2)))])
@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_fused_unpack_fft_single_rank(dims, perms, direction):
    """transpose!(…; fft=…) == numpy.fft along the new contiguous dim of the plain
    transpose! result, at the FFT's own tolerance (8 eps log2(L) max|X|)."""
    _, (ux, uy, uz) = _single_rank_chain(dims, torch.complex128, perms)
    ux.data.view(torch.float64).normal_()
    for dst, src in ((uy, ux), (uz, uy)):
        t = pa.Transposition(dst, src)
        pa.transpose_(t)
        torch.cuda.synchronize()
        plain = dst.data.cpu().numpy()           # torch (row-major) shape: contiguous dim LAST
        L = plain.shape[-1]
        fused = pa.PencilArray.undef(torch.complex128, dst.pencil)
        t2 = pa.Transposition(fused, src)
        if L < 8 or L & (L - 1):
            with pytest.raises(pa.ArgumentError):
                pa.transpose_(t2, fft=direction)
            continue
        n0 = pa.launch_count()
        pa.transpose_(t2, fft=direction)
        torch.cuda.synchronize()
        assert pa.launch_count() - n0 == 1       # ONE kernel: unpack and transform
        ref = np.fft.fft(plain, axis=-1) if direction == "forward" else np.fft.ifft(plain, axis=-1) * L
        tol = 8 * np.finfo(np.float64).eps * np.log2(L) * np.abs(ref).max()
        assert np.abs(fused.data.cpu().numpy() - ref).max() <= tol


def test_fused_fft_refusals():
    _, (ux, uy, uz) = _single_rank_chain((16, 32, 8), torch.complex128, ((2, 1, 3), (3, 2, 1)))
    with pytest.raises(pa.ArgumentError):      # aliased src / dst
        check(lib.pa_transpose(pa.Transposition(uy, ux).plan.h, None, ptr(ux.data), ptr(ux.data),
                               1 | pa._lib.PA_FFT_FORWARD, stream_ptr()))
    _, (fx, fy, fz) = _single_rank_chain((16, 32, 8), torch.float64, ((2, 1, 3), (3, 2, 1)))
    with pytest.raises(pa.ArgumentError):      # not ComplexF64
        pa.transpose_(pa.Transposition(fy, fx), fft="forward")
    # a line length that is not a power of two (21) is refused whatever the layout
    topo = pa.MPITopology(pa.COMM_SELF, (1, 1))
    px = pa.Pencil(topo, (21, 16, 8), (2, 3))
    a = pa.PencilArray.undef(torch.complex128, px)
    with pytest.raises(pa.ArgumentError):
        pa.fft_(a, "forward")


@pytest.mark.parametrize("dims", [(16, 8, 32), (64, 12, 8), (1024, 3, 5), (8, 8, 8)])
def test_in_place_line_fft_and_local_permute_fft(dims):
    """`fft_(u)`: the in-place transform along the contiguous dim (same kernel, linear
    gather); and a local permutation that keeps the contiguous dim, transformed on the way."""
    topo = pa.MPITopology(pa.COMM_SELF, (1, 1))
    px = pa.Pencil(topo, dims, (2, 3))
    u = pa.PencilArray.undef(torch.complex128, px)
    u.data.view(torch.float64).normal_()
    before = u.data.cpu().numpy()
    L = dims[0]
    n0 = pa.launch_count()
    pa.fft_(u, "forward")
    torch.cuda.synchronize()
    assert pa.launch_count() - n0 == 1
    ref = np.fft.fft(before, axis=-1)
    tol = 8 * np.finfo(np.float64).eps * np.log2(L) * np.abs(ref).max()
    assert np.abs(u.data.cpu().numpy() - ref).max() <= tol
    pa.fft_(u, "backward")
    torch.cuda.synchronize()
    assert np.abs(u.data.cpu().numpy() / L - before).max() <= tol
    # x stays contiguous, (y, z) swap places: permuted copy + transform in one kernel
    p2 = pa.Pencil(px, decomp_dims=(2, 3), permute=pa.Permutation(1, 3, 2))
    v = pa.PencilArray.undef(torch.complex128, p2)
    pa.transpose_(pa.Transposition(v, u), fft="forward")
    torch.cuda.synchronize()
    ref2 = np.fft.fft(u.data.cpu().numpy(), axis=-1).transpose(1, 0, 2)  # torch dims (z,y,x)->(y,z,x)
    assert np.abs(v.data.cpu().numpy() - ref2).max() <= 8 * np.finfo(np.float64).eps * np.log2(L) * np.abs(ref2).max()


@pytest.mark.parametrize("dims", [(16, 32, 8), (64, 64, 64), (8, 128, 16)])
def test_three_dimensional_fft_like_pencilffts(dims):
    """fft along x in place, then x->y and y->z transposes with the next transform fused:
    the z-pencil array, read in logical order, is numpy.fft.fftn of the input; the backward
    chain returns N^3 times the input."""
    _, (ux, uy, uz) = _single_rank_chain(dims, torch.complex128, ((2, 1, 3), (3, 2, 1)))
    ux.data.view(torch.float64).normal_()
    G = ux.logical().cpu().numpy().copy()
    orig = ux.data.clone()
    pa.fft_(ux, "forward")
    pa.transpose_(pa.Transposition(uy, ux), fft="forward")
    pa.transpose_(pa.Transposition(uz, uy), fft="forward")
    torch.cuda.synchronize()
    ref = np.fft.fftn(G)
    n = math.prod(dims)
    tol = 8 * np.finfo(np.float64).eps * np.log2(n) * np.abs(ref).max()
    assert np.abs(uz.logical().cpu().numpy() - ref).max() <= tol
    pa.fft_(uz, "backward")
    pa.transpose_(pa.Transposition(uy, uz), fft="backward")
    pa.transpose_(pa.Transposition(ux, uy), fft="backward")
    torch.cuda.synchronize()
    assert (ux.data / n - orig).abs().max().item() <= 8 * np.finfo(np.float64).eps * np.log2(n) * orig.abs().max().item()
