"""CPU, world_size 2 over gloo: the row-partitioned multi-GPU driver's host logic — nnz-balanced
split, slice rebasing, and the layout of the gathered result — with the ORACLE standing in for
the per-rank kernel (allowed in tests/ only).  The GPU path itself is covered by -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import random_csr, random_dense


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from basic_sparse_matrix_b200.gpu import partition_rows
        from oracle import ref_numpy
        rng = np.random.default_rng(11)          # same operands on every rank (B replicated)
        m, k, n = 301, 97, 6
        v, ci, ri = random_csr(rng, m, k, np.float64, giant_row=200, giant_len=900)
        b = random_dense(rng, k, n, np.float64)
        bounds = partition_rows(ri, world).astype(np.int64)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        # slice with rebased row pointer, exactly what bsm_csr_upload_rows sends to the device
        e0, e1 = int(ri[r0]), int(ri[r1])
        local = ref_numpy.mul_dense_rowmajor(v[e0:e1], ci[e0:e1], ri[r0:r1 + 1] - ri[r0], b)
        # all-gather(v): unequal blocks -> one broadcast per root into its slot of the full result
        full = torch.zeros((m, n), dtype=torch.float64)
        full[r0:r1] = torch.from_numpy(local)
        for root in range(world):
            s0, s1 = int(bounds[root]), int(bounds[root + 1])
            if s1 > s0:
                blk = full[s0:s1].contiguous()
                dist.broadcast(blk, src=root)
                full[s0:s1] = blk
        want = ref_numpy.mul_dense_rowmajor(v, ci, ri, b)
        ok = np.array_equal(full.numpy().view(np.uint64), want.view(np.uint64))
        nnz_share = (e1 - e0) / float(ri[-1])
        q.put((rank, ok, r0, r1, nnz_share))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_row_partition_and_gather_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == 301      # contiguous cover
    assert all(0.2 < r[4] < 0.8 for r in res)                                  # nnz-balanced
