"""PencilIO binary layout against the reference's definition
(src/PencilIO/mpi_io.jl; test/io.jl:56-85): the file holds the global array in the
pencil's MEMORY order (= PermutedDimsArray(gather(x), perm), column-major) or the ranks'
parent arrays in column-major grid order (chunks).

CPU part: the C planner's run table (`pa_io_sizes` / `pa_io_run_offset`) interpreted with
NumPy must reproduce those bytes for every rank of the reference's test decompositions.
GPU part: real files written from device arrays, re-read with another decomposition,
JSON sidecar keys, error cases."""
import ctypes as C
import itertools
import json
import math
import os

import numpy as np
import pytest

import pencilarrays_b200 as pa
from pencilarrays_b200._lib import lib, check, i64arr
from oracle import pencil_oracle as O
from util import make_ranks, perm_of

CONFIGS = [  # (grid, dims, decomp, perm, extra, dtype)
    ((2, 2), (16, 21, 41), (2, 3), None, (), np.float64),
    ((2, 3), (16, 21, 41), (1, 3), (2, 3, 1), (), np.float64),
    ((3, 2), (16, 21, 41), (1, 2), (3, 2, 1), (3,), np.float32),
    ((4,), (8, 6, 5), (2,), (3, 1, 2), (2, 2), np.complex128),
    ((5, 1), (3, 7, 4), (1, 3), (2, 1, 3), (), np.float64),      # ranks that own nothing
    ((1, 1), (6, 5, 4), (2, 3), None, (), np.int16),
]


def expected_file(G, perm, chunks, locals_, grid):
    """Reference layout of one dataset from the global (logical-order) array."""
    N = len(grid) + 0
    if not chunks:
        nd = G.ndim
        axes = tuple(range(nd)) if perm is None else tuple(p - 1 for p in perm) + tuple(range(len(perm), nd))
        return np.transpose(G, axes).reshape(-1, order="F").tobytes()
    # column-major linear order of the process-grid coordinates (mpi_io.jl:412-424)
    order = sorted(range(len(locals_)), key=lambda r: tuple(reversed(locals_[r][0])))
    return b"".join(locals_[r][1].reshape(-1, order="F").tobytes() for r in order)


def build(cfg):
    grid, dims, decomp, perm, extra, dt = cfg
    ranks = make_ranks(grid)
    rng = np.random.default_rng(3)
    shape = tuple(dims) + tuple(extra)
    if np.dtype(dt).kind == "c":
        G = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)
    elif np.dtype(dt).kind == "i":
        G = rng.integers(-30000, 30000, size=shape).astype(dt)
    else:
        G = rng.standard_normal(shape).astype(dt)
    pens, locals_ = [], []
    for er in ranks:
        p = pa.Pencil(er.topo, dims, decomp, permute=perm_of(perm))
        pens.append(p)
        rl = pa.range_local(p)  # logical, 1-based inclusive ranges
        box = tuple(slice(r.start - 1, r.stop - 1) for r in rl)
        loc = G[box + (slice(None),) * len(extra)]
        if perm is not None:
            loc = np.transpose(loc, tuple(q - 1 for q in perm) + tuple(range(len(dims), loc.ndim)))
        locals_.append((tuple(c - 1 for c in er.topo.coords_local), np.asfortranarray(loc)))
    return G, pens, locals_


@pytest.mark.parametrize("cfg", CONFIGS, ids=[str(c[:4]) for c in CONFIGS])
@pytest.mark.parametrize("chunks", [False, True])
def test_run_table_reproduces_reference_layout(cfg, chunks):
    grid, dims, decomp, perm, extra, dt = cfg
    G, pens, locals_ = build(cfg)
    es = np.dtype(dt).itemsize
    want = expected_file(G, perm, chunks, locals_, grid)
    got = bytearray(len(want))
    covered = 0
    for p, (_, loc) in zip(pens, locals_):
        gb, lb, nr, rb, fo = (C.c_int64() for _ in range(5))
        check(lib.pa_io_sizes(p._h, len(extra), i64arr(extra), es, int(chunks), C.byref(gb), C.byref(lb),
                              C.byref(nr), C.byref(rb), C.byref(fo)))
        assert gb.value == len(want) and lb.value == loc.size * es
        flat = loc.reshape(-1, order="F").tobytes()
        assert nr.value * rb.value == len(flat)
        for r in range(nr.value):
            off = C.c_int64()
            check(lib.pa_io_run_offset(p._h, len(extra), i64arr(extra), es, int(chunks), r, C.byref(off)))
            got[off.value:off.value + rb.value] = flat[r * rb.value:(r + 1) * rb.value]
            covered += rb.value
    assert covered == len(want)          # the ranks' runs tile the dataset exactly once
    assert bytes(got) == want


# ------------------------------------------------------------------------------- GPU
def _upload(p, loc, extra, tdt):
    import torch
    u = pa.PencilArray.undef(tdt, p, *extra)
    u.data.view(torch.uint8).reshape(-1).copy_(torch.from_numpy(
        np.frombuffer(loc.reshape(-1, order="F").tobytes(), dtype=np.uint8).copy()))
    return u


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS[:5], ids=[str(c[:4]) for c in CONFIGS[:5]])
def test_files_match_reference_layout_and_round_trip(cfg, tmp_path):
    import torch
    grid, dims, decomp, perm, extra, dt = cfg
    tdt = {np.float64: torch.float64, np.float32: torch.float32, np.complex128: torch.complex128}[dt]
    G, pens, locals_ = build(cfg)
    fname = str(tmp_path / "fields.bin")
    arrays = [_upload(p, loc, extra, tdt) for p, (_, loc) in zip(pens, locals_)]
    plus1 = [_upload(p, loc + 1, extra, tdt) for p, (_, loc) in zip(pens, locals_)]
    # every emulated rank writes its part of the same file (rank 0 first: it creates the file)
    files = [pa.open_(pa.MPIIODriver(), fname, er_p.topology.comm, write=True, create=(i == 0),
                      append=(i != 0)) for i, er_p in enumerate(pens)]
    for f in files:
        f.position = 0
    for f, a, b in zip(files, arrays, plus1):
        f.write("field_1", a, chunks=False)
        f.write("field_2", a, chunks=True)
        f["pair"] = (a, b)
    for f in reversed(files):
        f.close()                                   # rank 0 last: its sidecar is the one that stays
    raw = open(fname, "rb").read()
    n1 = len(expected_file(G, perm, False, locals_, grid))
    locals_b = [(c, l + 1) for c, l in locals_]
    assert raw[:n1] == expected_file(G, perm, False, locals_, grid)
    assert raw[n1:2 * n1] == expected_file(G, perm, True, locals_, grid)
    assert raw[2 * n1:3 * n1] == expected_file(G, perm, False, locals_, grid)
    assert raw[3 * n1:4 * n1] == expected_file(G + 1, perm, False, locals_b, grid)
    assert len(raw) == 4 * n1
    meta = json.load(open(fname + ".json"))
    assert meta["driver"] == {"type": "MPIIODriver", "version": "0.9.4"}
    d = meta["datasets"]["field_2"]
    jl = {np.float64: "Float64", np.float32: "Float32", np.complex128: "ComplexF64"}[dt]
    mem = tuple(dims) if perm is None else tuple(dims[q - 1] for q in perm)
    assert d == {"permutation": None if perm is None else list(perm), "extra_dims": list(extra),
                 "decomposed_dims": list(decomp), "process_dims": list(grid),
                 "julia_endian_bom": "0x04030201", "little_endian": True, "element_type": jl,
                 "dims_logical": list(dims) + list(extra), "dims_memory": list(mem) + list(extra),
                 "chunks": True, "offset_bytes": n1, "size_bytes": n1}
    assert meta["datasets"]["pair"]["dims_memory"] == list(mem) + list(extra) + [2]
    assert meta["datasets"]["pair"]["size_bytes"] == 2 * n1

    # read back: same decomposition (both layouts), then ANOTHER decomposition (discontiguous only)
    for r, p in enumerate(pens):
        with pa.open_(pa.MPIIODriver(), fname, p.topology.comm, read=True) as ff:
            for name in ("field_1", "field_2"):
                y = pa.PencilArray.undef(tdt, p, *extra)
                pa.read_(ff, y, name)
                assert y.data.cpu().numpy().tobytes() == arrays[r].data.cpu().numpy().tobytes()
            ya, yb = pa.PencilArray.undef(tdt, p, *extra), pa.PencilArray.undef(tdt, p, *extra)
            pa.read_(ff, (ya, yb), "pair")
            assert yb.data.cpu().numpy().tobytes() == plus1[r].data.cpu().numpy().tobytes()
            with pytest.raises(RuntimeError):
                pa.read_(ff, ya, "field not in file")
            yw = pa.PencilArray.undef(torch.int16 if tdt != torch.int16 else torch.float32, p, *extra)
            with pytest.raises(RuntimeError):
                pa.read_(ff, yw, "field_1")         # element type differs
    topo1 = pa.MPITopology(pa.COMM_SELF, (1,) * len(grid))
    p1 = pa.Pencil(topo1, dims, decomp, permute=perm_of(perm))   # the whole array on one rank
    with pa.open_(pa.MPIIODriver(), fname, pa.COMM_SELF, read=True) as ff:
        y = pa.PencilArray.undef(tdt, p1, *extra)
        pa.read_(ff, y, "field_1")
        nd = G.ndim
        axes = tuple(range(nd)) if perm is None else tuple(q - 1 for q in perm) + tuple(range(len(perm), nd))
        assert y.data.cpu().numpy().tobytes() == np.transpose(G, axes).reshape(-1, order="F").tobytes()
        if math.prod(grid) > 1:
            with pytest.raises(RuntimeError):       # chunks need the writing topology (:329-336)
                pa.read_(ff, y, "field_2")
    # file without metadata: first dataset at offset 0 (:261-276)
    nometa = str(tmp_path / "nometa.bin")
    os.symlink(fname, nometa)
    with pa.open_(pa.MPIIODriver(), nometa, pa.COMM_SELF, read=True) as ff:
        y = pa.PencilArray.undef(tdt, p1, *extra)
        with pytest.raises(pa.ArgumentError):
            pa.read_(ff, y, "name_doesnt_matter")
        pa.read_(ff, y)
        assert y.data.cpu().numpy().tobytes() == np.transpose(G, axes).reshape(-1, order="F").tobytes()


@pytest.mark.gpu
def test_large_runs_go_through_the_staging_pipeline(tmp_path):
    """> 32 MiB per run and many pieces: the double-buffered staging must not lose a byte."""
    import torch
    topo = pa.MPITopology(pa.COMM_SELF, (1, 1))
    p = pa.Pencil(topo, (256, 192, 160), (2, 3), permute=pa.Permutation(2, 3, 1))
    u = pa.PencilArray.undef(torch.float64, p)
    u.data.normal_()
    fname = str(tmp_path / "big.bin")
    with pa.open_(pa.MPIIODriver(), fname, pa.COMM_SELF, write=True, create=True) as ff:
        ff["u"] = u
    assert open(fname, "rb").read() == u.data.cpu().numpy().tobytes()
    v = pa.PencilArray.undef(torch.float64, p)
    with pa.open_(pa.MPIIODriver(), fname, pa.COMM_SELF, read=True) as ff:
        pa.read_(ff, v, "u")
    assert torch.equal(u.data, v.data)
