#!/usr/bin/env python3
"""tools/probe_near_diag.py — upper bound on what sharing the near-diagonal B rows between consecutive rows could
give the vector kernel on the headline matrix: time the 3-D Laplacian with its i-1 / i+1 entries REMOVED
(5 gathers per row instead of 7; B, C and the slice geometry unchanged). A row-block kernel that loads the
columns row-1 .. row+RB once per block of RB rows still issues (4 RB + RB + 2) / RB = 5.5 loads per row at RB = 4,
so it cannot beat this."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from basic_sparse_matrix_b200 import Csr, gen, gpu
    torch.cuda.set_device(0)
    gpu.init(0)
    stream = torch.cuda.Stream()
    gpu.set_stream(stream.cuda_stream)
    g = 256
    v, ci, ri, dims = gen.laplacian(g, g, g)
    rows = np.repeat(np.arange(g ** 3, dtype=np.int64), np.diff(ri.astype(np.int64)))
    keep = np.abs(ci.astype(np.int64) - rows) != 1
    counts = np.bincount(rows[keep], minlength=g ** 3)
    ri5 = np.zeros(g ** 3 + 1, np.uint64)
    np.cumsum(counts, out=ri5[1:])
    out = []
    with torch.cuda.stream(stream):
        for name, (vv, cc, rr) in (("7-point", (v, ci, ri)), ("5 entries (no i-1, i+1)", (v[keep], ci[keep], ri5))):
            A = gpu.DeviceCsr.from_host(Csr.from_raw_parts((g ** 3, g ** 3), vv, cc, rr))
            for n in (128, 64):
                B = gpu.DeviceDense.generate(g ** 3, n, seed=5, mode=gen.MODE_EXACT)
                C = gpu.DeviceDense.alloc(g ** 3, n)
                total_ms, per = bench.time_device_steps(torch, A, B, C, 10, 3, None)
                out.append({"matrix": name, "n": n, "nnz": int(rr[-1]), "ms": round(total_ms / 10, 4), "ms_best": round(min(per), 4),
                            "launch": gpu.last_launch_info()})
                print(json.dumps(out[-1]), flush=True)
                B.close()
                C.close()
            A.close()


if __name__ == "__main__":
    main()
