#!/usr/bin/env python
"""bench.py -- `transpose!` throughput (GiB/s moved), the metric of BASELINE.json.

One STEP = the x -> y -> z -> y -> x chain of four `transpose!` calls
(x<->y and y<->z, both directions, PencilFFTs' usual permutations
None -> (2,1,3) -> (3,2,1)) over a synthetic ComplexF64 grid.  Weak scaling:
every GPU holds 2 GiB of the array (512^3 ComplexF64 per GPU), so that
N = 8 is exactly BASELINE configs[3] (1024^3 ComplexF64, process grid (4,2)).
`value` = 4 * global_bytes / 2^30 / step_time, whole job, inputs resident in
HBM; `e2e` = the same with the step's input copied from pinned host memory and
its result copied back inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]        # this framework
  python bench.py --impl reference ...                       # CPU port of the reference path
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = float(1 << 30)
PER_GPU = (512, 512, 512)           # ComplexF64 shard per GPU: 2 GiB
CHAIN = [((2, 3), None), ((1, 3), (2, 1, 3)), ((1, 2), (3, 2, 1))]  # x, y, z pencils
LEGS = ["x->y", "y->z", "z->y", "y->x"]
# --workload: "cfg4" (default, above) or "cfg5" = BASELINE configs[4]: Float32,
# 1 GiB per GPU ((1024,512,512) per GPU -> 2048x1024x1024 at N = 8), perms None -> (2,3,1) -> (3,1,2)
WORKLOADS = {
    "cfg4": dict(per_gpu=PER_GPU, chain=CHAIN, dtype="complex128", itemsize=16, tname="ComplexF64",
                 perms="None->(2,1,3)->(3,2,1)", note="2 GiB per GPU; N=8 is BASELINE configs[3]"),
    "cfg5": dict(per_gpu=(1024, 512, 512),
                 chain=[((2, 3), None), ((1, 3), (2, 3, 1)), ((1, 2), (3, 1, 2))],
                 dtype="float32", itemsize=4, tname="Float32", perms="None->(2,3,1)->(3,1,2)",
                 note="1 GiB per GPU; N=8 is BASELINE configs[4]"),
}


def grid_and_dims(n, per_gpu=PER_GPU):
    grid = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[n]
    mult = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]
    return grid, tuple(a * b for a, b in zip(per_gpu, mult))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region: the sampler
    runs from before the warm-up (nvidia-smi needs ~1 s to start) at 10 ms
    period; rows are time-stamped and only those inside [mark_start, mark_stop]
    (host clock, bracketing the timed steps) are reported.  If the region is
    shorter than the sampling allows, the rows within 250 ms of it are used and
    `window` says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "10"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t = time.time()
            while not self.rows and time.time() - t < 3.0:  # wait for the first sample
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()  # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                c = [x.strip() for x in r.split(",")]
                if len(c) < 7:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for nme, v in zip(names, c[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            return sm, mx, reasons

        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        window = "timed region"
        if len(inside) < 2:
            inside = [r for r in self.rows if t0 - 0.25 <= r[0] <= t1 + 0.25]
            window = "timed region +-250 ms (region shorter than the sampling period)"
        sm, mx, reasons = parse(inside)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "region_ms": round((t1 - t0) * 1e3, 1),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------- reference arm
def step_with(cts, bufs):
    def f(nthreads):
        for i, ct in enumerate(cts):
            ct.run([b[i % 2] for b in bufs], [b[(i + 1) % 2] for b in bufs], nthreads=nthreads)
    return f


def pick_threads(step, cores, nranks):
    """The CPU arm may use every host thread; on a big shared box more threads is
    not always faster (OpenMP barriers), so the fastest of a few counts is kept."""
    cands = sorted({c for c in (nranks, 2 * nranks, 4 * nranks, 8, 16, 32, 64, cores)
                    if 1 <= c <= cores})
    best, best_t = cands[0], float("inf")
    for c in cands:
        step(c)
        t0 = time.perf_counter()
        step(c)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def run_reference(args):
    """CPU port of the reference path (oracle/pa_oracle.c) on the host cores:
    all N ranks emulated in one process, one worker per rank (+ spare threads
    split each rank's loops), exchange = memcpy.  Bounded sample: the same chain
    on a grid with every axis halved (1/8 of the arm's volume)."""
    import numpy as np
    from oracle import c_oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.gpus
    sample = tuple(s // 2 for s in PER_GPU)
    grid, dims = grid_and_dims(n, sample)
    nranks = math.prod(grid)
    cores = len(os.sched_getaffinity(0)) or 1
    dtype = np.complex128
    steps_cfg = CHAIN + [CHAIN[1], CHAIN[0]]
    cts = [c_oracle.CTranspose(grid, dims, steps_cfg[i][0], steps_cfg[i][1], steps_cfg[i + 1][0],
                               steps_cfg[i + 1][1], (), dtype) for i in range(4)]
    rng = np.random.default_rng(42)
    bufs = []
    for r in range(nranks):
        nmax = max(max(ct.sz[r][0], ct.sz[r][1]) for ct in cts)
        a = rng.standard_normal(2 * nmax).view(np.complex128)
        bufs.append([a, np.zeros(nmax, dtype=dtype)])
    orig = [b[0].copy() for b in bufs]

    def step():
        ph = [0.0, 0.0, 0.0]
        for i, ct in enumerate(cts):
            srcs = [b[i % 2] for b in bufs]
            dsts = [b[(i + 1) % 2] for b in bufs]
            p = ct.run(srcs, dsts, nthreads=cores)
            ph = [x + y for x, y in zip(ph, p)]
        return ph

    cores = pick_threads(step_with(cts, bufs), cores, nranks)

    def step():  # noqa: F811 -- same chain, with the thread count that ran fastest
        ph = [0.0, 0.0, 0.0]
        for i, ct in enumerate(cts):
            p = ct.run([b[i % 2] for b in bufs], [b[(i + 1) % 2] for b in bufs], nthreads=cores)
            ph = [x + y for x, y in zip(ph, p)]
        return ph

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    phases = [0.0, 0.0, 0.0]
    for _ in range(args.steps):
        phases = [x + y for x, y in zip(phases, step())]
    dt = (time.perf_counter() - t0) / args.steps
    ok = all(np.array_equal(o.view(np.uint8), b[0].view(np.uint8)) for o, b in zip(orig, bufs))
    gbytes = math.prod(dims) * 16
    val = 4 * gbytes / GIB / dt
    sample_txt = (f"x->y->z->y->x on a {dims[0]}x{dims[1]}x{dims[2]} ComplexF64 grid "
                  f"(every axis of the arm's grid halved), {nranks} emulated rank(s), grid {grid}")
    print(json.dumps({
        "impl": "reference", "metric": "transpose_GiB_per_s", "value": round(val, 3), "unit": "GiB/s",
        "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (bytes)",
        "data": "synthetic",
        "config": {"workload": workload_name(n), "sample": sample_txt, "round_trip_bit_exact": bool(ok)},
        "cpu_baseline": {"value": round(val, 3), "unit": "GiB/s", "cores": cores if cores < nranks else nranks * (cores // nranks),
                         "kind": "port", "sample": sample_txt,
                         "phase_s_per_step": [round(p / args.steps, 4) for p in phases]},
        "e2e": {"value": round(val, 3), "unit": "GiB/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_name(n, wl="cfg4"):
    W = WORKLOADS[wl]
    grid, dims = grid_and_dims(n, W["per_gpu"])
    return (f"x->y->z->y->x transpose! chain, {dims[0]}x{dims[1]}x{dims[2]} {W['tname']}, "
            f"process grid {grid}, perms {W['perms']} ({W['note']})")


# ------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pencilarrays_b200 as pa

    n = args.gpus
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == n, f"--gpus {n} but WORLD_SIZE={world}: launch with torch.distributed.run"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the transpose! path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = pa.comm_world() if n > 1 else pa.COMM_SELF
    rank = comm.rank
    W = WORKLOADS[args.workload]
    chain_cfg, isz = W["chain"], W["itemsize"]
    grid, dims = grid_and_dims(n, W["per_gpu"])
    topo = pa.MPITopology(comm, grid)
    px = pa.Pencil(topo, dims, chain_cfg[0][0])
    py = pa.Pencil(px, decomp_dims=chain_cfg[1][0], permute=pa.Permutation(*chain_cfg[1][1]))
    pz = pa.Pencil(py, decomp_dims=chain_cfg[2][0], permute=pa.Permutation(*chain_cfg[2][1]))
    dt = getattr(torch, W["dtype"])
    ux, uy, uz = (pa.PencilArray.undef(dt, p) for p in (px, py, pz))
    gen = torch.Generator(device="cuda").manual_seed(42 + rank)
    ux.data.view(torch.float64 if isz == 16 else dt).normal_(generator=gen)
    orig = ux.data.clone()
    methods = {"alltoallv": pa.Alltoallv(), "pointtopoint": pa.PointToPoint(),
               "peerput": pa.PeerPut(), "peerget": pa.PeerGet()}
    # auto: the one-sided put path over NVLink (fastest, profiles/r1_bench_n8_*.json); if the
    # CUDA-IPC window cannot be set up on ANY rank, every rank takes the NCCL PointToPoint path
    name = ("peerput" if n > 1 else "pointtopoint") if args.method == "auto" else args.method
    pairs = [(uy, ux), (uz, uy), (uy, uz), (ux, uy)]
    while True:
        method = methods[name]
        ok_here = 1
        try:
            ts = [pa.Transposition(d, s, method=method) for d, s in pairs]
        except pa.PencilError as e:  # e.g. IPC not permitted in this container
            ok_here, ts = 0, None
            print(f"[rank {rank}] {name} unavailable: {e}", file=sys.stderr)
        if n > 1:
            flag = torch.tensor([ok_here], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok_here = int(flag.item())
        if ok_here:
            break
        if args.method != "auto" or name == "pointtopoint":
            raise SystemExit(f"method {name} could not be set up")
        name = "pointtopoint"
    overlap = not args.no_overlap
    if args.remote_ctas is not None:
        pa.check(pa.lib.pa_set_tunable(b"remote_ctas", args.remote_ctas))
    if args.nccl_fences:
        pa.check(pa.lib.pa_set_tunable(b"nccl_fences", 1))
    if args.no_nccl_register:
        pa.check(pa.lib.pa_set_tunable(b"nccl_register", 0))

    def barrier():
        torch.cuda.synchronize()
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if n == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def chain(evs=None):
        for i, t in enumerate(ts):
            pa.transpose_(t, waitall=True, overlap=overlap)
            if evs is not None:
                evs[i + 1].record()

    # ---- device-resident timing -------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        chain()
    leg_ms = [0.0] * 4
    barrier()
    sampler.mark_start()
    n0 = pa.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    legs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    e0.record()
    for k in range(args.steps):
        legs[k][0].record()
        chain(legs[k])
    e1.record()
    barrier()
    sampler.mark_stop()
    launches = pa.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    for k in range(args.steps):
        for i in range(4):
            leg_ms[i] += legs[k][i].elapsed_time(legs[k][i + 1]) / args.steps
    leg_ms = [max_over_ranks(x) for x in leg_ms]
    ok = bool(torch.equal(ux.data.view(torch.uint8), orig.view(torch.uint8)))
    gbytes = math.prod(dims) * isz
    value = 4 * gbytes / GIB / (ms * 1e-3)

    # ---- placement check at full size (not timed): a round trip alone would also pass for a
    # wrong-but-invertible shuffle, so every rank fills x with a function of the GLOBAL logical
    # index and checks that after x->y and y->z each element sits where the pencil geometry says:
    # parent(u)[perm * I] == global[I + offset]  (arrays.jl:327-337)
    def pattern(pen):
        rl = pa.range_local(pen)
        ax = [torch.arange(r.start - 1, r.stop - 1, device="cuda", dtype=torch.float64) for r in rl]
        lin = ax[0][:, None, None] + dims[0] * (ax[1][None, :, None] + dims[1] * ax[2][None, None, :])
        if isz == 16:
            return torch.complex(lin, -lin)
        return torch.remainder(lin, 16777216.0).to(dt)  # exact in Float32

    good, perr = 0, None
    try:
        ux.logical().copy_(pattern(px))
        pa.transpose_(ts[0], waitall=True, overlap=overlap)
        pa.transpose_(ts[1], waitall=True, overlap=overlap)
        good = int(bool(torch.equal(uy.logical(), pattern(py))) and
                   bool(torch.equal(uz.logical(), pattern(pz))))
    except Exception as e:  # never lose the bench line over the checker
        perr = f"checker error: {type(e).__name__}: {e}"[:200]
    flag = torch.tensor([good], device="cuda")
    if n > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank takes part, whatever happened above
    placement = perr if perr else bool(flag.item())
    torch.cuda.empty_cache()

    # ---- per-kernel roofline (kernels timed alone on the current stream) -------
    peak, peak_src = measured_peak()
    shard = ux.data.numel() * isz
    kern = {}
    if n == 1:
        # each leg IS one launch of the fused permuting kernel (K3); live numbers from the timed steps
        for name, m in zip(LEGS, leg_ms):
            kern[f"K3 fused permute {name}"] = {"ms": round(m, 4), "alg_bytes": 2 * shard,
                                                "GBps": round(2 * shard / m / 1e6, 1)}
    else:
        from pencilarrays_b200._lib import lib, check, PA_STAGE_SELF
        import ctypes as C
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for name, t, src, dst in (("x->y", ts[0], ux, uy), ("y->z", ts[1], uy, uz)):
            info = t.plan.info
            if info.dim == 0:
                continue
            check(lib.pa_pencil_reserve(t.Po._h, max(1, info.send_bytes), max(1, info.recv_bytes)))
            sp, _, rp, _ = t.Po.buffers()
            for op, label in ((0, "K1 pack"), (1, "K2 unpack")):
                def run():
                    for p in range(1, info.nproc + 1):
                        peer = t.plan.peer(p)
                        if op == 0:
                            check(lib.pa_pack(t.plan.h, p, C.c_void_p(src.data_ptr()),
                                              C.c_void_p(rp if peer.is_self else sp), st))
                        else:
                            check(lib.pa_unpack(t.plan.h, p, C.c_void_p(rp), C.c_void_p(dst.data_ptr()), st))
                for _ in range(3):
                    run()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    run()
                b.record()
                torch.cuda.synchronize()
                m = a.elapsed_time(b) / 5
                nb = 2 * (info.length_in if op == 0 else info.length_out) * isz
                kern[f"{label} {name} (all {info.nproc} blocks)"] = {
                    "ms": round(m, 4), "alg_bytes": nb, "GBps": round(nb / m / 1e6, 1)}
        chain()  # restore a consistent state after the isolated kernels scribbled on uy/uz
    dom = min(kern.items(), key=lambda kv: kv[1]["GBps"]) if kern else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("bytes_per_launch")
    except Exception:
        pass

    # ---- BASELINE configs[1] beside it (N == 1): 256^3 Float64, 1 GPU, pack/unpack kernel only ----
    cfg1 = None
    if n == 1 and args.workload == "cfg4":
        cfg1 = {}
        topo1 = pa.MPITopology(pa.COMM_SELF, (1, 1))
        q1 = pa.Pencil(topo1, (256, 256, 256), (2, 3))
        a1 = pa.PencilArray.undef(torch.float64, q1)
        a1.data.normal_()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        for perm in ((2, 1, 3), (2, 3, 1), (3, 2, 1), (3, 1, 2), (1, 3, 2), None):
            q2 = pa.Pencil(q1, decomp_dims=(1, 3),
                           permute=pa.NoPermutation() if perm is None else pa.Permutation(*perm))
            b1 = pa.PencilArray.undef(torch.float64, q2)
            t1 = pa.Transposition(b1, a1)
            tot = 0.0
            for it in range(8):
                flush.zero_()  # evict src/dst from L2 between launches
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                pa.transpose_(t1)
                e_.record()
                torch.cuda.synchronize()
                if it >= 3:
                    tot += s_.elapsed_time(e_) / 5
            nb = 2 * a1.data.numel() * 8
            cfg1[f"x->y perm {perm}"] = {"ms": round(tot, 4), "GBps": round(nb / tot / 1e6, 1),
                                         "frac_of_hbm_peak": round(nb / tot / 1e6 / peak, 3)}
        cfg1["note"] = ("256^3 Float64 x->y on 1 GPU = one fused K3 launch per transpose!; L2 flushed "
                        "(256 MiB memset) before every timed launch; 2*s*n algorithmic bytes")

    # ---- end to end: host buffers in, host buffers out -------------------------
    hin = torch.empty(ux.data.shape, dtype=dt).pin_memory()
    hin.copy_(orig)
    hout = torch.empty(ux.data.shape, dtype=dt).pin_memory()
    e2e_steps = max(4, min(args.steps, 8))
    # Every step copies its input from pinned host memory and its result back.  PCIe is full
    # duplex, so the steps are software-pipelined over two x-pencil buffers: the device -> host
    # copy of step i runs on its own stream while step i+1 uploads and transposes.
    ux2 = pa.PencilArray.undef(dt, px)
    ts2 = [pa.Transposition(uy, ux2, method=method), ts[1], ts[2],
           pa.Transposition(ux2, uy, method=method)]
    sets = [(ux, ts), (ux2, ts2)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    cur = torch.cuda.current_stream()

    def e2e_step(i):
        u, tl = sets[i % 2]
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_out[i % 2])  # this buffer's previous result has left the device
            u.data.copy_(hin, non_blocking=True)
            ev_in[i % 2].record(s_in)
        cur.wait_event(ev_in[i % 2])
        for t in tl:
            pa.transpose_(t, waitall=True, overlap=overlap)
        ev_done[i % 2].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[i % 2])
            hout.copy_(u.data, non_blocking=True)
            ev_out[i % 2].record(s_out)

    def e2e_drain():
        cur.wait_event(ev_out[0])
        cur.wait_event(ev_out[1])

    e2e_step(0)
    e2e_step(1)
    e2e_drain()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(e2e_steps):
        e2e_step(i)
    e2e_drain()
    b.record()
    barrier()
    e2e_ms = max_over_ranks(a.elapsed_time(b)) / e2e_steps
    e2e_ok = bool(torch.equal(hout.view(torch.uint8), hin.view(torch.uint8)))
    e2e_val = 4 * gbytes / GIB / (e2e_ms * 1e-3)

    # ---- exchange timing (N > 1): library CUDA-event sections, sequential phases ----
    sections = None
    if n > 1:
        sections = {}
        for name, t in zip(LEGS[:2], ts[:2]):
            if t.dim is None:
                continue
            t.enable_timing(True)
            barrier()  # ranks enter together: otherwise the sections contain the skew
            pa.transpose_(t, waitall=True, overlap=False)
            tm = t.timings()
            t.enable_timing(False)
            info = t.plan.info
            sections[name] = {"pack_ms": round(tm.pack_ms, 3), "exchange_ms": round(tm.exchange_ms, 3),
                              "unpack_ms": round(tm.unpack_ms, 3), "total_ms": round(tm.total_ms, 3),
                              "send_bytes": info.send_bytes,
                              "nvlink_GBps_out": round(info.send_bytes / max(tm.exchange_ms, 1e-6) / 1e6, 1),
                              "nvlink_frac_of_770": round(info.send_bytes / max(tm.exchange_ms, 1e-6) / 1e6 / 770, 3)}
        chain()

    # ---- CPU baseline beside it (rank 0, N == 1 only) -------------------------------
    cpu = None
    if n == 1 and rank == 0 and not args.no_cpu:
        cpu = cpu_baseline()

    if rank == 0:
        out = {
            "metric": "transpose_GiB_per_s", "value": round(value, 2), "unit": "GiB/s", "n_gpus": n,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("c128" if isz == 16 else "f32") + " (bytes; pure data movement)", "data": "synthetic",
            "config": {"workload": workload_name(n, args.workload), "method": repr(method), "overlap": overlap,
                       "l2": "inputs (2 GiB per GPU) far larger than the 126 MB L2; no flush needed",
                       "round_trip_bit_exact": ok, "placement_exact_full_size": placement, "leg_ms": dict(zip(LEGS, [round(x, 4) for x in leg_ms]))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": round(e2e_val, 2), "unit": "GiB/s", "ms_per_step": round(e2e_ms, 3),
                    "h2d_bytes_per_step": int(shard), "d2h_bytes_per_step": int(shard),
                    "steps": e2e_steps, "round_trip_bit_exact": e2e_ok,
                    "note": "per step: pinned host -> device copy of the x-pencil array, the four "
                            "transposes, device -> pinned host copy of the result (bytes per GPU); "
                            "steps are software-pipelined over two device buffers so the result "
                            "download of step i overlaps the upload + transposes of step i+1 "
                            "(PCIe is full duplex); timed from first upload to last download"},
            "roofline": None if dom is None else {
                "bound": "hbm", "kernel": dom[0], "achieved": dom[1]["GBps"], "peak": peak,
                "unit": "GB/s", "frac": round(dom[1]["GBps"] / peak, 4), "traffic": traffic,
                "peak_source": peak_src,
                "alg_bytes_per_launch": dom[1]["alg_bytes"],
                "timing": ("CUDA events on the launching stream inside the timed steps" if n == 1 else
                           "CUDA events, kernels launched alone on the current stream in this run")},
            "kernels": kern,
            "configs1_256cubed_f64": cfg1,
            "sections": sections,
            "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline():
    """oracle/pa_oracle.c timed on the host cores on a bounded sample (about 10-20 s)."""
    import numpy as np
    from oracle import c_oracle
    dims = (256, 256, 256)
    grid = (1, 1)
    cores = len(os.sched_getaffinity(0)) or 1
    cfg = CHAIN + [CHAIN[1], CHAIN[0]]
    cts = [c_oracle.CTranspose(grid, dims, cfg[i][0], cfg[i][1], cfg[i + 1][0], cfg[i + 1][1], (),
                               np.complex128) for i in range(4)]
    nel = math.prod(dims)
    a = np.random.default_rng(1).standard_normal(2 * nel).view(np.complex128)
    b = np.zeros(nel, dtype=np.complex128)
    bufs = [a, b]
    cores = pick_threads(step_with(cts, [bufs]), cores, 1)

    def step():
        for i, ct in enumerate(cts):
            ct.run([bufs[i % 2]], [bufs[(i + 1) % 2]], nthreads=cores)

    step()
    t0 = time.perf_counter()
    reps = 0
    while True:
        step()
        reps += 1
        el = time.perf_counter() - t0
        if el > 10.0 or reps >= 200:
            break
    val = 4 * nel * 16 / GIB / (el / reps)
    return {"value": round(val, 3), "unit": "GiB/s", "cores": cores, "kind": "port",
            "sample": f"x->y->z->y->x chain on a 256^3 ComplexF64 grid, 1 emulated rank, "
                      f"{reps} repetitions in {el:.1f} s (oracle/pa_oracle.c, OpenMP over {cores} threads)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="auto",
                    choices=["auto", "pointtopoint", "alltoallv", "peerput", "peerget"])
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-nccl-register", action="store_true",
                    help="staged methods: plain cudaMalloc arenas, no NCCL user-buffer registration")
    ap.add_argument("--nccl-fences", action="store_true",
                    help="one-sided methods: fence with NCCL groups instead of NVLink flags")
    ap.add_argument("--remote-ctas", type=int, default=None,
                    help="grid cap of the PeerPut/PeerGet kernels (tunable remote_ctas)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.gpus not in (1, 2, 4, 8):
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
