#!/usr/bin/env python3
"""tools/probe_box_stencil.py — 27-point box stencil on a g^3 grid (host-built, uploaded): does the line detection
(distances nx-1, nx, nx+1 from the diagonal -> nx) give the vector kernel its line-aligned rows per warp?
Prints the launch geometry and time for n = 64 f64, default and with the rows per warp forced to nx - 1."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def box_stencil(g):
    n = g ** 3
    i = np.arange(n, dtype=np.int64)
    x, y, z = i % g, (i // g) % g, i // (g * g)
    cols, rows = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = (x + dx >= 0) & (x + dx < g) & (y + dy >= 0) & (y + dy < g) & (z + dz >= 0) & (z + dz < g)
                rows.append(i[ok])
                cols.append((i + dx + dy * g + dz * g * g)[ok])
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    ri = np.zeros(n + 1, np.uint64)
    np.cumsum(np.bincount(rows, minlength=n), out=ri[1:])
    vals = np.where(rows == cols, 26.0, -1.0)
    return vals, cols.astype(np.uint64), ri


def main():
    import torch
    from basic_sparse_matrix_b200 import Csr, gen, gpu
    torch.cuda.set_device(0)
    gpu.init(0)
    stream = torch.cuda.Stream()
    gpu.set_stream(stream.cuda_stream)
    g, n = 128, 64
    v, ci, ri = box_stencil(g)
    with torch.cuda.stream(stream):
        A = gpu.DeviceCsr.from_host(Csr.from_raw_parts((g ** 3, g ** 3), v, ci, ri))
        B = gpu.DeviceDense.generate(g ** 3, n, seed=5, mode=gen.MODE_EXACT)
        C = gpu.DeviceDense.alloc(g ** 3, n)
        for name, kw in (("default", {}), ("rows_per_warp=127", dict(rows_per_warp=g - 1)), ("rows_per_warp=128", dict(rows_per_warp=g))):
            t = gpu.make_tuning("vector", **kw)
            total_ms, per = bench.time_device_steps(torch, A, B, C, 10, 3, t)
            print(json.dumps({"point": name, "nnz": int(ri[-1]), "ms": round(total_ms / 10, 4), "ms_best": round(min(per), 4),
                              "launch": gpu.last_launch_info()}), flush=True)


if __name__ == "__main__":
    main()
