"""The C++ planner (libpa_b200: geometry, wire layout, kernel descriptors)
against the oracle, on CPU: the descriptors are interpreted with NumPy
(tests/util.apply_block) and every buffer must match the oracle bit for bit --
send_buf / recv_buf wire layout (Transpositions.jl:380-416) included."""
import math

import numpy as np
import pytest

import pencilarrays_b200 as pa
from pencilarrays_b200.transpositions import _Plan
from oracle import pencil_oracle as O
from util import CASES, DTYPES, beq, build_chain, emulate_transpose_with_descriptors


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("method", [pa.PointToPoint(), pa.Alltoallv()], ids=["p2p", "a2av"])
def test_chain_matches_oracle(case, method):
    dtype = DTYPES[case["it"]]
    extra = case["extra"]
    ranks, steps = build_chain(case)
    g = O.global_pattern(case["dims"], extra, case["it"])
    cur_o = O.scatter(g, [po for (_, po) in steps[0]], extra, dtype)
    G = O.gather(cur_o)
    cur = [a.data.reshape(-1, order="F").copy() for a in cur_o]
    for k in range(1, len(steps)):
        pin, pout = steps[k - 1], steps[k]
        # oracle
        nxt_o = [O.OArray.undef(dtype, po, *extra) for (_, po) in pout]
        states = O.transpose_all(nxt_o, cur_o, keep_states=True)
        assert beq(O.gather(nxt_o), G)  # the reference's own check (test/transpose.jl:6-22)
        # C planner, descriptors interpreted by NumPy
        plans = [_Plan(pin[r][0], pout[r][0], extra, case["it"], method) for r in range(len(ranks))]
        nxt = [np.zeros(max(1, a.data.size), dtype=dtype) for a in nxt_o]
        sends, recvs = emulate_transpose_with_descriptors(plans, cur, nxt, dtype)
        for r in range(len(ranks)):
            assert beq(nxt[r][:nxt_o[r].data.size], nxt_o[r].data.reshape(-1, order="F")), (k, r)
            info = plans[r].info
            assert info.length_in == cur_o[r].data.size and info.length_out == nxt_o[r].data.size
            if states is not None:  # wire layout
                st = states[r]
                ns, nr = info.send_bytes // case["it"], info.recv_bytes // case["it"]
                assert beq(sends[r][:ns], st.send_buf[:ns]), ("send_buf", k, r)
                assert beq(recvs[r][:nr], st.recv_buf[:nr]), ("recv_buf", k, r)
                assert info.dim == O.transposition_dim(pin[r][1], pout[r][1])
            else:
                assert info.dim == 0
        cur_o, cur = nxt_o, [a[:o.data.size].copy() for a, o in zip(nxt, nxt_o)]


def test_fused_self_block_equals_pack_then_unpack():
    """K3 (src -> dest in one pass) must equal pack-to-recv_buf + unpack for the self block."""
    case = CASES[0]
    dtype = DTYPES[case["it"]]
    ranks, steps = build_chain(case)
    g = O.global_pattern(case["dims"], (), case["it"])
    src_o = O.scatter(g, [po for (_, po) in steps[0]], (), dtype)
    for r in range(len(ranks)):
        plan = _Plan(steps[0][r][0], steps[1][r][0], (), case["it"], pa.PointToPoint())
        n_out = plan.info.length_out
        src = src_o[r].data.reshape(-1, order="F")
        from util import apply_block
        a = np.zeros(n_out, dtype=dtype)
        b = np.zeros(n_out, dtype=dtype)
        recv = np.zeros(plan.info.recv_bytes // case["it"], dtype=dtype)
        me = plan.info.self_index
        apply_block(plan.block(0, me), src, recv)
        apply_block(plan.block(1, me), recv, a)
        apply_block(plan.block(2), src, b)
        assert beq(a, b)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_put_descriptors_match_oracle(case):
    """PeerPut: every block written straight into the DESTINATION rank's parent
    array (descriptor op 3, in the peer's layout) + the fused self block must
    reproduce the oracle's result without any staging buffer."""
    from util import apply_block
    dtype, extra = DTYPES[case["it"]], case["extra"]
    ranks, steps = build_chain(case)
    g = O.global_pattern(case["dims"], extra, case["it"])
    cur_o = O.scatter(g, [po for (_, po) in steps[0]], extra, dtype)
    for k in range(1, len(steps)):
        nxt_o = [O.OArray.undef(dtype, po, *extra) for (_, po) in steps[k]]
        O.transpose_all(nxt_o, cur_o)
        plans = [_Plan(steps[k - 1][r][0], steps[k][r][0], extra, case["it"], pa.PeerPut())
                 for r in range(len(ranks))]
        if plans[0].info.dim != 0:
            cur = [a.data.reshape(-1, order="F") for a in cur_o]
            nxt = [np.zeros(max(1, a.data.size), dtype=dtype) for a in nxt_o]
            for r, pl in enumerate(plans):
                apply_block(pl.block(2), cur[r], nxt[r])
                for p in range(1, pl.info.nproc + 1):
                    peer = pl.peer(p)
                    if not peer.is_self:
                        apply_block(pl.block(3, p), cur[r], nxt[peer.world_rank])
            for r, a in enumerate(nxt_o):
                assert beq(nxt[r][:a.data.size], a.data.reshape(-1, order="F")), (k, r)
            # PeerGet: the same result pulled out of the SOURCE ranks' parents (op 4)
            nxt = [np.zeros(max(1, a.data.size), dtype=dtype) for a in nxt_o]
            for r, pl in enumerate(plans):
                apply_block(pl.block(2), cur[r], nxt[r])
                for p in range(1, pl.info.nproc + 1):
                    peer = pl.peer(p)
                    if not peer.is_self:
                        apply_block(pl.block(4, p), cur[peer.world_rank], nxt[r])
            for r, a in enumerate(nxt_o):
                assert beq(nxt[r][:a.data.size], a.data.reshape(-1, order="F")), ("get", k, r)
        cur_o = nxt_o


def test_counts_are_symmetric():
    """send count r->q equals recv count q<-r for every pair (what Isend/Irecv rely on)."""
    for case in CASES:
        ranks, steps = build_chain(case)
        for k in range(1, len(steps)):
            plans = [_Plan(steps[k - 1][r][0], steps[k][r][0], case["extra"], case["it"],
                           pa.PointToPoint()) for r in range(len(ranks))]
            if plans[0].info.dim == 0:
                continue
            for r, pl in enumerate(plans):
                tot_s = tot_r = 0
                for p in range(1, pl.info.nproc + 1):
                    peer = pl.peer(p)
                    if peer.is_self:
                        assert peer.world_rank == r
                        assert peer.recv_offset + peer.recv_count == pl.info.recv_bytes
                        continue
                    tot_s += peer.send_count
                    tot_r += peer.recv_count
                    q = plans[peer.world_rank]
                    back = [q.peer(pp) for pp in range(1, q.info.nproc + 1)]
                    back = [b for b in back if b.world_rank == r][0]
                    assert back.recv_count == peer.send_count
                    assert back.send_count == peer.recv_count
                assert tot_s == pl.info.send_bytes
                assert tot_r + pl.info.length_self * case["it"] == pl.info.recv_bytes


@pytest.mark.parametrize("itemsize", [24, 12, 48])
def test_non_power_of_two_element_sizes(itemsize):
    """Elements such as SVector{3,Float64} (24 B) or 3 x Float32 (12 B) move as
    several power-of-two words: the planner adds an innermost word dimension."""
    word = max(w for w in (1, 2, 4, 8, 16) if itemsize % w == 0)  # the planner's word size
    wdt = np.dtype((np.void, word))
    edt = np.dtype((np.void, itemsize))
    case = dict(name="odd_elsize", grid=(2, 2), dims=(6, 7, 5), extra=(2,), it=itemsize,
                chain=[((2, 3), None), ((1, 3), (2, 3, 1)), ((1, 2), (3, 2, 1)), ((1, 2), (1, 3, 2))])
    ranks, steps = build_chain(case)
    rng = np.random.default_rng(7)
    n_glob = math.prod(case["dims"]) * math.prod(case["extra"])
    gbytes = rng.integers(0, 256, size=(n_glob, itemsize), dtype=np.uint8)
    g = gbytes.reshape(case["dims"] + case["extra"] + (itemsize,), order="F")
    cur_o = O.scatter(g, [po for (_, po) in steps[0]], case["extra"], edt)
    for k in range(1, len(steps)):
        nxt_o = [O.OArray.undef(edt, po, *case["extra"]) for (_, po) in steps[k]]
        O.transpose_all(nxt_o, cur_o)
        plans = [_Plan(steps[k - 1][r][0], steps[k][r][0], case["extra"], itemsize, pa.PointToPoint())
                 for r in range(len(ranks))]
        cur = [np.ascontiguousarray(a.data.reshape(-1, order="F")).view(wdt) for a in cur_o]
        nxt = [np.zeros(max(1, a.data.size * itemsize // word), dtype=wdt) for a in nxt_o]
        emulate_transpose_with_descriptors(plans, cur, nxt, wdt)
        for r, a in enumerate(nxt_o):
            want = np.ascontiguousarray(a.data.reshape(-1, order="F")).view(np.uint8)
            assert nxt[r].view(np.uint8)[:want.size].tobytes() == want.tobytes(), (k, r)
        cur_o = nxt_o


@pytest.mark.parametrize("case", [c for c in CASES if math.prod(c["grid"]) > 1][:10],
                         ids=[c["name"] for c in CASES if math.prod(c["grid"]) > 1][:10])
@pytest.mark.parametrize("nparts", [1, 2, 3, 7])
def test_chunked_blocks_tile_the_wire_layout(case, nparts):
    """The chunked PointToPoint schedule (tunable p2p_chunks): the pack chunks of a
    block, run one after the other, fill exactly the bytes the whole pack fills; each
    chunk is ONE contiguous piece of the wire block; the receiver cuts its unpack at
    the same wire offsets (sender and receiver never talk about the cuts)."""
    import ctypes as C
    from pencilarrays_b200._lib import lib, check, BlockDesc
    from util import apply_block
    dtype, it, extra = DTYPES[case["it"]], case["it"], case["extra"]
    ranks, steps = build_chain(case)
    g = O.global_pattern(case["dims"], extra, it)
    cur_o = O.scatter(g, [po for (_, po) in steps[0]], extra, dtype)
    for k in range(1, min(3, len(steps))):
        nxt_o = [O.OArray.undef(dtype, po, *extra) for (_, po) in steps[k]]
        states = O.transpose_all(nxt_o, cur_o, keep_states=True)
        plans = [_Plan(steps[k - 1][r][0], steps[k][r][0], extra, it, pa.PointToPoint())
                 for r in range(len(ranks))]
        if plans[0].info.dim != 0:
            for r, pl in enumerate(plans):
                src = cur_o[r].data.reshape(-1, order="F").copy()
                send = np.zeros(max(1, pl.info.send_bytes // it), dtype=dtype)
                recv = states[r].recv_buf.copy()
                dst = np.zeros(max(1, pl.info.length_out), dtype=dtype)
                for p in range(1, pl.info.nproc + 1):
                    peer = pl.peer(p)
                    if peer.is_self:
                        apply_block(pl.block(1, p), recv, dst)
                        continue
                    end_s, end_r = peer.send_offset, peer.recv_offset
                    for c in range(nparts):
                        d, off, nb = BlockDesc(), C.c_int64(), C.c_int64()
                        check(lib.pa_plan_get_chunk(pl.h, 0, p, c, nparts, C.byref(d), C.byref(off),
                                                    C.byref(nb)))
                        if nb.value:
                            assert off.value == end_s  # contiguous, in order
                        end_s += nb.value
                        apply_block(d, src, send)
                        d2, off2, nb2 = BlockDesc(), C.c_int64(), C.c_int64()
                        check(lib.pa_plan_get_chunk(pl.h, 1, p, c, nparts, C.byref(d2), C.byref(off2),
                                                    C.byref(nb2)))
                        if nb2.value:
                            assert off2.value == end_r
                        end_r += nb2.value
                        apply_block(d2, recv, dst)
                        # the matching chunk on the sending side of this receive has the same size
                        q = peer.world_rank
                        back = [pp for pp in range(1, pl.info.nproc + 1)
                                if plans[q].peer(pp).world_rank == pl.peer(pl.info.self_index).world_rank][0]
                        d3, off3, nb3 = BlockDesc(), C.c_int64(), C.c_int64()
                        check(lib.pa_plan_get_chunk(plans[q].h, 0, back, c, nparts, C.byref(d3),
                                                    C.byref(off3), C.byref(nb3)))
                        assert nb3.value == nb2.value
                        # ... and that sender aims at exactly this receive slot of mine
                        assert plans[q].peer(back).remote_recv_offset == peer.recv_offset
                    assert end_s == peer.send_offset + peer.send_count
                    assert end_r == peer.recv_offset + peer.recv_count
                ns = pl.info.send_bytes // it
                assert beq(send[:ns], states[r].send_buf[:ns])
                assert beq(dst[:nxt_o[r].data.size], nxt_o[r].data.reshape(-1, order="F"))
        cur_o = nxt_o
