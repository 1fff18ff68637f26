#!/usr/bin/env python
"""BASELINE configs[1] under ncu: 256^3 Float64, x -> y with destination permutation
(3,2,1) (the slowest of the six), rotating over 6 array pairs like bench.py does.
  ncu --set full --clock-control none -k regex:k_box -s 12 -c 4 python profiles/prof_cfg1.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pencilarrays_b200 as pa  # noqa: E402

topo = pa.MPITopology(pa.COMM_SELF, (1, 1))
q1 = pa.Pencil(topo, (256, 256, 256), (2, 3))
q2 = pa.Pencil(q1, decomp_dims=(1, 3), permute=pa.Permutation(3, 2, 1))
srcs = [pa.PencilArray.undef(torch.float64, q1) for _ in range(6)]
dsts = [pa.PencilArray.undef(torch.float64, q2) for _ in range(6)]
for a in srcs:
    a.data.normal_()
ts = [pa.Transposition(d, s) for d, s in zip(dsts, srcs)]
for _ in range(4):
    for t in ts:
        pa.transpose_(t)
torch.cuda.synchronize()
