"""Randomised (seeded) pencil configurations: process grids, global sizes with
uneven / empty blocks, every permutation pair, extra dims, element sizes.
CPU: the C++ planner's descriptors (staged, put and get forms) interpreted by
NumPy vs the oracle.  GPU: the same through the kernels with emulated ranks."""
import itertools
import math
import random

import numpy as np
import pytest

import pencilarrays_b200 as pa
from pencilarrays_b200.transpositions import _Plan
from oracle import pencil_oracle as O
from util import DTYPES, apply_block, beq, build_chain, emulate_transpose_with_descriptors


def random_case(rng: random.Random, idx: int):
    N = rng.choice([2, 3, 3, 3, 4])
    M = rng.randint(1, N - 1)
    grid = tuple(rng.choice([1, 2, 2, 3, 4]) for _ in range(M))
    while math.prod(grid) > 12:
        grid = tuple(max(1, g - 1) for g in grid)
    dims = tuple(rng.choice([1, 2, 3, 5, 8, 12, 16, 21]) for _ in range(N))
    it = rng.choice([2, 4, 8, 16])
    extra = tuple(rng.choice([2, 3]) for _ in range(rng.choice([0, 0, 1, 2]))) if N < 4 else ()
    perms = [None] + [p for p in itertools.permutations(range(1, N + 1)) if p != tuple(range(1, N + 1))]
    decomp = tuple(rng.sample(range(1, N + 1), M))
    chain = [(decomp, rng.choice(perms))]
    for _ in range(rng.randint(2, 3)):
        d = list(chain[-1][0])
        if rng.random() < 0.8:  # change exactly one decomposed dimension (or none: local permute)
            i = rng.randrange(M)
            free = [x for x in range(1, N + 1) if x not in d]
            if free:
                d[i] = rng.choice(free)
        chain.append((tuple(d), rng.choice(perms)))
    return dict(name=f"rand{idx}", grid=grid, dims=dims, extra=extra, it=it, chain=chain)


RNG = random.Random(20260922)
RANDOM_CASES = [random_case(RNG, i) for i in range(60)]


def _run(case, on_gpu):
    dtype, extra, it = DTYPES[case["it"]], case["extra"], case["it"]
    ranks, steps = build_chain(case)
    g = O.global_pattern(case["dims"], extra, it)
    cur_o = O.scatter(g, [po for (_, po) in steps[0]], extra, dtype)
    G = O.gather(cur_o)
    for k in range(1, len(steps)):
        nxt_o = [O.OArray.undef(dtype, po, *extra) for (_, po) in steps[k]]
        O.transpose_all(nxt_o, cur_o)
        assert beq(O.gather(nxt_o), G)
        plans = [_Plan(steps[k - 1][r][0], steps[k][r][0], extra, it, pa.PeerPut())
                 for r in range(len(ranks))]
        cur = [np.ascontiguousarray(a.data.reshape(-1, order="F")) for a in cur_o]
        sizes = [a.data.size for a in nxt_o]
        if not on_gpu:
            staged = [np.zeros(max(1, n), dtype=dtype) for n in sizes]
            emulate_transpose_with_descriptors(plans, cur, staged, dtype)
            results = [("staged", staged)]
            if plans[0].info.dim != 0:
                for op, label in ((3, "put"), (4, "get")):
                    out = [np.zeros(max(1, n), dtype=dtype) for n in sizes]
                    for r, pl in enumerate(plans):
                        apply_block(pl.block(2), cur[r], out[r])
                        for p in range(1, pl.info.nproc + 1):
                            peer = pl.peer(p)
                            if peer.is_self:
                                continue
                            if op == 3:
                                apply_block(pl.block(3, p), cur[r], out[peer.world_rank])
                            else:
                                apply_block(pl.block(4, p), cur[peer.world_rank], out[r])
                    results.append((label, out))
            for label, out in results:
                for r, a in enumerate(nxt_o):
                    assert beq(out[r][:a.data.size], a.data.reshape(-1, order="F")), (case, k, r, label)
        else:
            import torch
            from gpu_util import dev_bytes, host_bytes, emulate_transpose_gpu
            dcur = [dev_bytes(c) if c.size else torch.zeros(1, dtype=torch.uint8, device="cuda") for c in cur]
            dnxt = [torch.full((max(1, n * it),), 0x77, dtype=torch.uint8, device="cuda") for n in sizes]
            emulate_transpose_gpu(plans, dcur, dnxt, fused_self=bool(k % 2))
            torch.cuda.synchronize()
            for r, a in enumerate(nxt_o):
                want = np.ascontiguousarray(a.data.reshape(-1, order="F")).view(np.uint8)
                assert host_bytes(dnxt[r])[:want.size].tobytes() == want.tobytes(), (case, k, r)
        cur_o = nxt_o


@pytest.mark.parametrize("case", RANDOM_CASES[:25], ids=[c["name"] for c in RANDOM_CASES[:25]])
def test_random_geometry_api_vs_oracle(case):
    """`range_local`, `range_remote`, `size_local`, `to_local` of the host mirror
    (computed by libpa_b200) against the oracle's axes for every rank."""
    import itertools as it_
    ranks, steps = build_chain(case)
    for pens in steps:
        for r, (p, op) in enumerate(pens):
            assert pa.range_local(p) == op.axes_local
            assert pa.range_local(p, pa.MemoryOrder()) == op.axes_local_perm
            assert pa.size_local(p, pa.MemoryOrder()) == op.size_local(True)
            assert pa.size_global(p) == op.size_global
            assert p.topology.coords_local == op.topology.coords_local
            for coords in it_.product(*[range(1, d + 1) for d in case["grid"]]):
                assert pa.range_remote(p, coords) == op.axes_all[coords]
                assert p.topology.rank_of(coords) == op.topology.rank_of(coords)
            probe = tuple(range(rg.start, min(rg.stop, rg.start + 2)) for rg in op.axes_local)
            for mem in (False, True):
                order = pa.MemoryOrder() if mem else pa.LogicalOrder()
                assert pa.to_local(p, probe, order) == op.to_local(probe, memory_order=mem)


@pytest.mark.parametrize("case", RANDOM_CASES, ids=[c["name"] for c in RANDOM_CASES])
def test_random_planner_vs_oracle(case):
    _run(case, on_gpu=False)


@pytest.mark.gpu
@pytest.mark.parametrize("case", RANDOM_CASES[:30], ids=[c["name"] for c in RANDOM_CASES[:30]])
def test_random_kernels_vs_oracle(case):
    _run(case, on_gpu=True)
