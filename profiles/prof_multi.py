#!/usr/bin/env python
"""The one-launch multi-peer put kernel (k_multi) with the peers emulated by local
arrays on ONE GPU: rank 0 of the (4,2) grid of BASELINE configs[3] (1024^3 ComplexF64)
stores its three remote x->y blocks (512 MiB each) and its one remote y->z block (1 GiB)
into stand-ins for the peers' `dest` arrays.  Gives the kernel's HBM-side ceiling (the
NVLink side needs real peers: bench.py --gpus 8) and something ncu can profile.

  python profiles/prof_multi.py [--cap CTAS]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pencilarrays_b200 as pa  # noqa: E402
from pencilarrays_b200._lib import lib, check  # noqa: E402
from pencilarrays_b200.transpositions import _Plan  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cap", type=int, default=0, help="grid cap (0 = tunable remote_ctas)")
    args = ap.parse_args()
    comm = pa.Comm(0, 8)
    topo = pa.MPITopology(comm, (4, 2))
    dims = (1024, 1024, 1024)
    px = pa.Pencil(topo, dims, (2, 3))
    py = pa.Pencil(px, decomp_dims=(1, 3), permute=pa.Permutation(2, 1, 3))
    pz = pa.Pencil(py, decomp_dims=(1, 2), permute=pa.Permutation(3, 2, 1))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.pa_set_device(torch.cuda.current_device()))
    for name, pi, po in (("x->y", px, py), ("y->z", py, pz)):
        plan = _Plan(pi, po, (), 16, pa.PeerPut())
        info = plan.info
        src = torch.empty(info.length_in * 16, dtype=torch.uint8, device="cuda").random_()
        peers = [torch.empty(info.length_out * 16, dtype=torch.uint8, device="cuda")
                 for _ in range(info.nproc)]
        arr = (C.c_void_p * info.nproc)(*[p.data_ptr() for p in peers])
        for cap in ([args.cap] if args.cap else [0, -8, 1 << 30]):
            def run():
                check(lib.pa_put_all(plan.h, C.c_void_p(src.data_ptr()), arr, cap, st))
            for _ in range(2):
                run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                run()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            nb = 2 * info.send_bytes
            print(json.dumps({"leg": name, "cap": cap, "remote_blocks": info.nproc - 1, "ms": round(ms, 4),
                              "alg_bytes": nb, "GBps": round(nb / ms / 1e6, 1)}), flush=True)
        del src, peers
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
