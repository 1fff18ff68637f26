"""Generates tests/golden/*.npz from the oracle (the reference cannot run in
this image -- no Julia, no MPI -- so fixtures are oracle outputs frozen at
commit time; see oracle/pencil_oracle.py for the parity status).
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pencil_oracle as O  # noqa: E402
from util import CASES, DTYPES  # noqa: E402

PICK = ["ref_transpose_2x2", "ref_unsorted", "ref_extra_dims", "ref_slab", "baseline_cfg1",
        "two_ranks_2x1", "empty_blocks"]

for case in CASES:
    if case["name"] not in PICK:
        continue
    dtype, extra = DTYPES[case["it"]], case["extra"]
    out = dict(grid=np.array(case["grid"]), dims=np.array(case["dims"]),
               extra=np.array(extra, dtype=np.int64), itemsize=np.array(case["it"]),
               nsteps=np.array(len(case["chain"])))
    g = O.global_pattern(case["dims"], extra, case["it"])
    pens = [O.make_pencils(case["grid"], case["dims"], d, p) for (d, p) in case["chain"]]
    for k, (d, p) in enumerate(case["chain"]):
        out[f"decomp{k}"] = np.array(d)
        out[f"perm{k}"] = np.array(p if p else (), dtype=np.int64)
    cur = O.scatter(g, pens[0], extra, dtype)
    for k in range(1, len(pens)):
        nxt = [O.OArray.undef(dtype, p, *extra) for p in pens[k]]
        O.transpose_all(nxt, cur)
        for r, a in enumerate(nxt):
            out[f"step{k}_rank{r}"] = np.ascontiguousarray(a.data.reshape(-1, order="F")).view(np.uint8)
        cur = nxt
    np.savez_compressed(os.path.join(HERE, case["name"] + ".npz"), **out)
    print("wrote", case["name"])
