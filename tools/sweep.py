#!/usr/bin/env python3
"""tools/sweep.py — time bsm_spmm_tuned over a list of tuning points on ONE resident workload.

    python tools/sweep.py --workload laplace3d_256_n128_f64 --steps 10 \
        --points "col_tile=64;reg_flavour=5,rows_per_slice=32;..."  [--algo vector] [--slice p/N]

Each point is `k=v,k=v` over the fields of bsm_tuning (include/bsm.h); an empty point is the
library heuristics. Prints one JSON line per point (ms, GFLOP/s, effective GB/s, fraction of the
measured HBM roofline, launch geometry) and checks that every point produces the same sampled
output rows as the first one (bitwise)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def sample(dense, gpu, ids):
    i = dense.info()
    s = np.dtype(i["dtype"]).itemsize
    out = np.empty((len(ids), i["cols"]), i["dtype"])
    for j, r in enumerate(ids):
        view = gpu.DeviceDense.borrow(i["ptr"] + int(r) * i["ld"] * s, 1, i["cols"], i["ld"], i["dtype"])
        out[j] = view.to_rowmajor()[0]
        view.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--points", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--slice", default="", help="p/N: time rank p's nnz-balanced row block of an N-way partition on this one GPU")
    args = ap.parse_args()

    import torch
    from basic_sparse_matrix_b200 import gen, gpu
    torch.cuda.set_device(0)
    gpu.init(0)
    stream = torch.cuda.Stream()
    gpu.set_stream(stream.cuda_stream)
    peak, _ = bench.load_peaks()
    kind, prm, n, dt = bench.WORKLOADS[args.workload]
    dtype = bench.NP_DTYPE[dt]
    s = np.dtype(dtype).itemsize
    if args.slice:
        pnum, nparts = (int(x) for x in args.slice.split("/"))
        bounds = gpu.partition_rows(bench.host_row_index(kind, prm), nparts).astype(np.int64)
        A = bench.make_device_csr(gpu, kind, prm, dtype, int(bounds[pnum]), int(bounds[pnum + 1]))
    else:
        A = bench.make_device_csr(gpu, kind, prm, dtype)
    ai = A.info()
    B = gpu.DeviceDense.generate(ai["cols"], n, seed=5, mode=gen.MODE_EXACT, dtype=dtype)
    C = gpu.DeviceDense.alloc(ai["rows"], n, dtype)
    bm = bench.bytes_min(ai["rows"], ai["nnz"], ai["cols"], n, s)
    ids = np.unique(np.concatenate([np.random.default_rng(1).integers(0, ai["rows"], 64), [0, ai["rows"] - 1]]))
    ref = None
    lines = []
    with torch.cuda.stream(stream):
        for pt in args.points.split(";"):
            kw = {k: int(v, 0) for k, v in (kv.split("=") for kv in pt.split(",") if kv)}
            tuning = gpu.make_tuning(kw.pop("algo", args.algo), **kw)
            try:
                total_ms, per = bench.time_device_steps(torch, A, B, C, args.steps, args.warmup, tuning)
            except Exception as ex:
                lines.append({"point": pt, "error": str(ex)[:200]})
                print(json.dumps(lines[-1]), flush=True)
                continue
            info = gpu.last_launch_info()
            got = sample(C, gpu, ids)
            if ref is None:
                ref = got
            t = total_ms / args.steps * 1e-3
            line = {"point": pt, "ms": round(total_ms / args.steps, 4), "ms_best": round(min(per), 4),
                    "gflops": round(2.0 * ai["nnz"] * n / t / 1e9, 1), "eff_gbs": round(bm / t / 1e9, 1),
                    "frac": round(bm / t / 1e9 / peak, 4), "same_as_first": bool(np.array_equal(got.view(np.uint8), ref.view(np.uint8))),
                    "launch": info}
            lines.append(line)
            print(json.dumps(line), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
