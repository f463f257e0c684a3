#!/usr/bin/env python3
"""bench.py — the reference's headline metric on B200: SpMM GFLOP/s + effective HBM GB/s
(% of the measured roofline) for Csr x Dense, next to the reference's CPU path on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3]): 3-D 7-point Laplacian on a 256^3 grid (16 777 216 rows,
117 047 296 nnz) x dense 16 777 216 x 128, f64, row-partitioned over the N GPUs with nnz-balanced
splits and B replicated (strong scaling: the problem is fixed, each rank owns a row block, no
data-path collective).  A "step" is one pass of the hot path (one `Csr::mul_dense`) over the
matrix.  `value` times the device-resident path with CUDA events.  `e2e` times the LITERAL reference
call — host Csr x host Dense -> zero-dropped host Csr (sparse.rs:426-446 incl. the result construction) —
through the C ABI from pinned host buffers (`bsm_mul_dense_host_into_*`: A up, B up in row chunks,
SpMM + result construction per row block, values / usize columns / row_index down); `e2e_dense` is the
same product returned as host Dense columns.  Operands are far larger than L2 (126 MB), so consecutive
steps cannot hit in cache (no explicit flush needed).

Every line carries `parity`: rows of this run's OWN device result (every rank's block; with the gathered
result also rows owned by other ranks) recomputed by the CPU oracle and compared bit for bit, and
`north_star_target`: the same matrix x 64 columns (>= 100 M nnz x 64, f64) in the same run.

One JSON line is printed by rank 0.  `--impl reference` times the reference's own CPU
implementation (the C restatement in oracle/, single-threaded like the Rust original) on the
host cores for the same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

def _baseline_metric():
    """BASELINE.json's metric string, verbatim (value = its GFLOP/s part, 2*nnz*ncols / t; the effective GB/s
    and roofline fraction it also names are the `effective_gbs` / `roofline` keys of the same line)."""
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "SpMM GFLOP/s + effective HBM GB/s (% roofline) at 1/2/4/8 B200 vs Rust CPU"


METRIC = _baseline_metric()

WORKLOADS = {
    # name: (kind, params, n, dtype)
    "laplace3d_256_n128_f64": ("laplace3d", dict(g=256), 128, "f64"),     # BASELINE configs[3] — the headline
    "laplace3d_256_n64_f64": ("laplace3d", dict(g=256), 64, "f64"),       # north_star target case (>=100M nnz x 64)
    "laplace2d_2048_n1_f64": ("laplace2d", dict(g=2048), 1, "f64"),       # configs[1] SpMV
    "rmat20_n64_f64": ("rmat", dict(scale=20, edges=100 << 20), 64, "f64"),   # configs[2]
    "rmat20_n64_f32": ("rmat", dict(scale=20, edges=100 << 20), 64, "f32"),
    "band_1m_hb32_n32_f32": ("band", dict(n=1 << 20, hb=32), 32, "f32"),  # configs[4] GPU side
    "band_1m_hb32_n1_f32": ("band", dict(n=1 << 20, hb=32), 1, "f32"),
    "band_1m_hb32_n64_f64": ("band", dict(n=1 << 20, hb=32), 64, "f64"),   # sweep-only: wider / f64 products on the band
    "band_1m_hb32_n128_f32": ("band", dict(n=1 << 20, hb=32), 128, "f32"),
    # sweep-only shapes (tools/sweep.py): 256-byte and 128-byte output rows on the stencil matrix
    "laplace3d_256_n32_f64": ("laplace3d", dict(g=256), 32, "f64"),
    "laplace3d_256_n64_f32": ("laplace3d", dict(g=256), 64, "f32"),
    "laplace3d_256_n16_f64": ("laplace3d", dict(g=256), 16, "f64"),
    "laplace3d_256_n128_f32": ("laplace3d", dict(g=256), 128, "f32"),
    "laplace3d_256_n8_f64": ("laplace3d", dict(g=256), 8, "f64"),
    "laplace3d_256_n1_f64": ("laplace3d", dict(g=256), 1, "f64"),     # SpMV on the headline matrix
    "laplace3d_256_n4_f64": ("laplace3d", dict(g=256), 4, "f64"),
    "laplace2d_4096_n64_f64": ("laplace2d", dict(g=4096), 64, "f64"),   # 5 entries per row, line length 4096
    "laplace3d_252_n128_f64": ("laplace3d", dict(g=252), 128, "f64"),   # line lengths that are not powers of two
    "laplace3d_250_n128_f64": ("laplace3d", dict(g=250), 128, "f64"),
    "laplace3d_252_n64_f64": ("laplace3d", dict(g=252), 64, "f64"),
    # uniformly random columns, Poisson(8) rows: regular row lengths without any locality in B
    "uniform22_n64_f64": ("rmat", dict(scale=22, edges=8 << 22, a=0.25, b=0.25, c=0.25), 64, "f64"),
    "uniform22_n32_f64": ("rmat", dict(scale=22, edges=8 << 22, a=0.25, b=0.25, c=0.25), 32, "f64"),
}
DEFAULT_WORKLOAD = "laplace3d_256_n128_f64"
NP_DTYPE = {"f64": np.float64, "f32": np.float32}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def bytes_min(rows, nnz, k_ref, n, s):
    """Algorithmic bytes per launch: values + u32 col_idx, u32 row_ptr, the B rows
    actually referenced, C written once."""
    return nnz * (s + 4) + (rows + 1) * 4 + k_ref * n * s + rows * n * s


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for nm in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
            v = getattr(nv, "nvmlClocksEventReason" + nm, None) or getattr(nv, "nvmlClocksThrottleReason" + nm, None)
            if v is not None:
                names[v] = nm
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(1.0)
        snake = {"HwSlowdown": "hw_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown",
                 "SwThermalSlowdown": "sw_thermal_slowdown", "SwPowerCap": "sw_power_cap",
                 "HwPowerBrakeSlowdown": "hw_power_brake_slowdown"}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(snake[r] for r in self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# workload construction (device-side generators; host side regenerates from the same hash)
# ---------------------------------------------------------------------------------------------
def host_row_index(kind, prm):
    """Row pointer of the FULL matrix in the reference layout (usize), for the nnz-balanced split."""
    from basic_sparse_matrix_b200 import gen
    if kind == "laplace3d":
        counts = gen.laplacian_row_counts(prm["g"], prm["g"], prm["g"])
    elif kind == "laplace2d":
        counts = gen.laplacian_row_counts(prm["g"], prm["g"], 1)
    elif kind == "band":
        i = np.arange(prm["n"], dtype=np.int64)
        counts = np.minimum(i, prm["hb"]) + 1 + np.minimum(prm["n"] - 1 - i, prm["hb"])
    else:
        return None
    ri = np.zeros(len(counts) + 1, dtype=np.uint64)
    np.cumsum(counts, out=ri[1:])
    return ri


def make_device_csr(gpu, kind, prm, dtype, r0=0, r1=None):
    if kind == "laplace3d":
        return gpu.DeviceCsr.laplacian(prm["g"], prm["g"], prm["g"], r0, r1, dtype)
    if kind == "laplace2d":
        return gpu.DeviceCsr.laplacian(prm["g"], prm["g"], 1, r0, r1, dtype)
    if kind == "band":
        return gpu.DeviceCsr.band(prm["n"], prm["hb"], r0, r1, dtype)
    if kind == "rmat":
        from basic_sparse_matrix_b200 import gen
        return gpu.DeviceCsr.rmat(prm["scale"], prm["edges"], a=prm.get("a", 0.57), b=prm.get("b", 0.19), c=prm.get("c", 0.19),
                                  seed=3, mode=gen.MODE_EXACT, dtype=dtype)
    raise ValueError(kind)


def k_ref_of(kind, prm, rows_total, r0, r1):
    """Distinct B rows referenced by rows [r0,r1)."""
    if kind == "laplace3d":
        reach = prm["g"] * prm["g"]
    elif kind == "laplace2d":
        reach = prm["g"]
    elif kind == "band":
        reach = prm["hb"]
    else:
        return rows_total
    return min(rows_total, r1 + reach) - max(0, r0 - reach)


ALGO_NAME = {1: "vector", 2: "merge", 3: "rowblock"}
KERNEL_NAME = {1: "spmm_rows_kernel", 2: "spmm_merge_kernel", 3: "spmm_rowblock_kernel"}
B_SEED = 5


def bind_to_gpu_numa_node(torch, local):
    """Pinned host buffers are first-touched by the allocating thread: run this process on the CPUs of the NUMA node the
    GPU hangs off, so that every rank's H2D / D2H traffic stays on its own socket. Returns what was found."""
    info = {"node": None, "cpus": None, "bound": False}
    try:
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip())
        cpus_txt = open(base + "/local_cpulist").read().strip()
        info.update(node=node, cpus=cpus_txt, pci=bus)
        cpus = set()
        for part in cpus_txt.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
        info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
    except Exception as ex:   # no sysfs / single node: nothing to bind
        info["note"] = str(ex)[:120]
    return info


def parity_row_ids(r0, r1, count=64, reach=0, rows_total=None):
    """Rows of the block [r0, r1) to recompute: its first and last rows, the rows one stencil reach inside it, and hashed rows."""
    from basic_sparse_matrix_b200 import gen
    nrows = r1 - r0
    if nrows <= 0:
        return np.zeros(0, np.int64)
    ids = {r0, r0 + 1, r1 - 2, r1 - 1, r0 + nrows // 2}
    if reach:
        ids.update([r0 + reach - 1, r0 + reach, r1 - reach - 1, r1 - reach])
    h = gen.hash_u64(977, np.arange(count, dtype=np.uint64) + np.uint64(r0))
    ids.update(int(r0 + int(x) % nrows) for x in h)
    return np.array(sorted(i for i in ids if r0 <= i < r1), dtype=np.int64)


def oracle_rows(kind, prm, n, dtype, row_ids, a_host=None):
    """The reference's value of the given GLOBAL rows of A x B (sequential sum in stored order, mul then add), recomputed
    by the CPU oracle from the counter-based generators: no operand travels."""
    from basic_sparse_matrix_b200 import gen
    from oracle import ref_numpy
    out = np.empty((len(row_ids), n), dtype)
    for j, r in enumerate(row_ids):
        r = int(r)
        if kind == "laplace3d":
            rv, rc, rr, _ = gen.laplacian(prm["g"], prm["g"], prm["g"], r, r + 1, dtype)
        elif kind == "laplace2d":
            rv, rc, rr, _ = gen.laplacian(prm["g"], prm["g"], 1, r, r + 1, dtype)
        elif kind == "band":
            rv, rc, rr, _ = gen.band(prm["n"], prm["hb"], r, r + 1, dtype)
        else:   # R-MAT: rows of the downloaded matrix
            v, ci, ri = a_host
            s, e = int(ri[r]), int(ri[r + 1])
            rv, rc, rr = v[s:e], ci[s:e], np.array([0, e - s], np.uint64)
        k_total = {"laplace3d": lambda: prm["g"] ** 3, "laplace2d": lambda: prm["g"] ** 2, "band": lambda: prm["n"]}.get(kind, lambda: 1 << prm["scale"])()
        uniq, inv = np.unique(rc.astype(np.int64), return_inverse=True)
        bsub = gen.dense_rows(k_total, n, B_SEED, gen.MODE_EXACT, 0.0, dtype, row_ids=uniq)
        out[j] = ref_numpy.mul_dense_rowmajor(rv, inv.astype(np.uint64), rr, bsub)[0] if len(rc) else 0
    return out


def bitwise_equal(a, b):
    a = np.where(a == 0, 0.0, a).astype(a.dtype)   # -0.0 == 0.0 to the reference (both dropped by insert)
    b = np.where(b == 0, 0.0, b).astype(b.dtype)
    return bool(np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8)))


def csr_rows_dense(v, ci, ri, row_ids, r_base, n, dtype):
    """Densify rows of a result Csr (reference layout) — zeros where insert dropped the value."""
    out = np.zeros((len(row_ids), n), dtype)
    for j, r in enumerate(row_ids):
        s, e = int(ri[r - r_base]), int(ri[r - r_base + 1])
        out[j, ci[s:e].astype(np.int64)] = v[s:e]
    return out


def sample_dense_rows(gpu, dense, row_ids):
    """Download selected rows of a device-resident dense matrix (one-row borrowed views)."""
    i = dense.info()
    sz = np.dtype(i["dtype"]).itemsize
    out = np.empty((len(row_ids), i["cols"]), i["dtype"])
    for j, r in enumerate(row_ids):
        view = gpu.DeviceDense.borrow(i["ptr"] + int(r) * i["ld"] * sz, 1, i["cols"], i["ld"], i["dtype"])
        out[j] = view.to_rowmajor()[0]
        view.close()
    return out


def time_device_steps(torch, A, B, C, steps, warmup, tuning=None, barrier=None):
    """W untimed + K timed launches on torch's current stream; per-launch CUDA events."""
    for _ in range(warmup):
        A.mul_dense(B, out=C, tuning=tuning)
    torch.cuda.synchronize()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        A.mul_dense(B, out=C, tuning=tuning)
        evs[i + 1].record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return total_ms, per


def run_config1(torch, gpu, peak, e=900_000):
    """BASELINE configs[0]: the reference's own `sd_mul` bench shape (benches/sparse_dense_mul.rs:6-35,
    f64 for u32): 1000x1000 Csr built by `e` unordered inserts x Dense 1000x10. Small enough for the CPU
    port to run IN FULL beside the GPU: device-resident kernel time, the literal host call
    (`Csr.mul_dense` -> zero-dropped Csr, uploads and download included) and the CPU restatement."""
    from basic_sparse_matrix_b200 import gen
    from oracle.ref_cpu import OracleCsr
    a, x = gen.bench_as_written(e)
    v, ci, ri = a.raw_parts()
    A = gpu.DeviceCsr.from_host(a)
    B = gpu.DeviceDense.from_host(x)
    C = gpu.DeviceDense.alloc(1000, 10, np.float64)
    total_ms, per = time_device_steps(torch, A, B, C, 20, 3)
    info = gpu.last_launch_info()
    t0 = time.perf_counter()
    for _ in range(5):
        out = a.mul_dense(x)
    t_host = (time.perf_counter() - t0) / 5
    o = OracleCsr.from_raw((1000, 1000), v, ci, ri)
    cols = [np.ascontiguousarray(c) for c in x.data]
    t_cpu = min(o.time_mul_dense_rows(cols, 0, 1000, faithful=True)[0] for _ in range(3))
    ref = o.mul_dense(cols, faithful=False)
    same = bool(np.array_equal(out.v, ref.v) and np.array_equal(out.col_index, ref.col_index) and np.array_equal(out.row_index, ref.row_index))
    nnz = int(ri[-1])
    for h in (A, B, C):
        h.close()
    return {"workload": f"bench_as_written_e{e}_n10_f64", "rows": 1000, "nnz": nnz, "n": 10, "dtype": "f64",
            "algo": ALGO_NAME.get(info["algo"], "?"), "ms_per_step": round(total_ms / 20, 4), "ms_best": round(min(per), 4),
            "gflops": round(2.0 * nnz * 10 / (total_ms / 20 * 1e-3) / 1e9, 2), "kernels_per_step": info["kernels"],
            "host_call_ms": round(t_host * 1e3, 3), "host_call_gflops": round(2.0 * nnz * 10 / t_host / 1e9, 3),
            "cpu_port_ms_full": round(t_cpu * 1e3, 3), "cpu_port_gflops": round(2.0 * nnz * 10 / t_cpu / 1e9, 4),
            "result_csr_equals_cpu_port": same,
            "note": "launch-bound on the GPU (11 MB of operands); the one config the CPU port runs in full"}


def run_extra(torch, gpu, name, steps, warmup, peak, tuning=None, label=None):
    """Secondary single-GPU workloads (kernel-only numbers, reported under other_workloads)."""
    from basic_sparse_matrix_b200 import gen
    kind, prm, n, dt = WORKLOADS[name]
    dtype = NP_DTYPE[dt]
    s = np.dtype(dtype).itemsize
    A = make_device_csr(gpu, kind, prm, dtype)
    ai = A.info()
    B = gpu.DeviceDense.generate(ai["cols"], n, seed=4, mode=gen.MODE_EXACT, dtype=dtype)
    C = gpu.DeviceDense.alloc(ai["rows"], n, dtype)
    total_ms, per = time_device_steps(torch, A, B, C, steps, warmup, tuning)
    info = gpu.last_launch_info()
    t = total_ms / steps * 1e-3
    bm = bytes_min(ai["rows"], ai["nnz"], ai["cols"], n, s)
    out = {"workload": label or name, "rows": ai["rows"], "nnz": ai["nnz"], "n": n, "dtype": dt,
           "algo": ALGO_NAME.get(info["algo"], "?"), "ms_per_step": round(total_ms / steps, 4),
           "ms_best": round(min(per), 4), "gflops": round(2.0 * ai["nnz"] * n / t / 1e9, 1),
           "eff_gbs": round(bm / t / 1e9, 1), "roofline_frac": round(bm / t / 1e9 / peak, 4),
           "kernels_per_step": info["kernels"]}
    for h in (A, B, C):
        h.close()
    return out


def run_solve(torch, gpu, n_rows=1 << 20, hb=32, nrhs=32):
    """BASELINE configs[4], the part after the factorisation: L y = b and L* x = y (lib.rs:28-65) for 32 right-hand sides on
    the device, against the CPU port of the same two loops (oracle/ref_solve_band.c, 1 thread), on a synthetic band factor
    (values do not matter for the timing; the factorisation itself stays on the CPU and is not timed). X is compared bit for bit."""
    from basic_sparse_matrix_b200 import Csr, gen
    from oracle import ref_solve
    f32 = np.float32
    rng = np.random.default_rng(11)
    l_band = np.zeros((n_rows, hb + 1), f32)                       # band storage of a lower-triangular factor
    l_band[:, :hb] = rng.uniform(-0.02, 0.02, (n_rows, hb)).astype(f32)
    l_band[:, hb] = rng.uniform(1.0, 2.0, n_rows).astype(f32)
    for i in range(min(hb, n_rows)):                                # columns < 0 do not exist
        l_band[i, :hb - i] = 0
    b_cols = gen.dense_rows(n_rows, nrhs, 6, gen.MODE_REAL, 0.5, f32).T.copy()
    L = gpu.DeviceCsr.from_host(Csr.from_raw_parts((n_rows, n_rows), *ref_solve.band_to_csr_lower(l_band)))
    Ls = gpu.DeviceCsr.from_host(Csr.from_raw_parts((n_rows, n_rows), *ref_solve.band_to_csr_upper(l_band)))
    B = gpu.DeviceDense.generate(n_rows, nrhs, seed=6, mode=gen.MODE_REAL, offset=0.5, dtype=f32)
    Y, X = gpu.DeviceDense.alloc(n_rows, nrhs, f32), gpu.DeviceDense.alloc(n_rows, nrhs, f32)
    L.forward_substitution(B, out=Y)                                # warm-up
    Ls.backward_substitution(Y, out=X)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    L.forward_substitution(B, out=Y)
    e[1].record()
    Ls.backward_substitution(Y, out=X)
    e[2].record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y_ref = np.stack([ref_solve.forward(l_band, b_cols[c]) for c in range(nrhs)])
    t1 = time.perf_counter()
    x_ref = np.stack([ref_solve.backward(l_band, y_ref[c]) for c in range(nrhs)])
    t2 = time.perf_counter()
    same = bitwise_equal(X.to_rowmajor().T, x_ref)
    nnz = int(L.info()["nnz"])
    for h in (L, Ls, B, Y, X):
        h.close()
    return {"workload": f"band_{n_rows}_hb{hb}_solve_n{nrhs}_f32", "what": "forward + backward substitution (lib.rs:28-65), factor given", "rows": n_rows,
            "factor_nnz": nnz, "forward_ms": round(e[0].elapsed_time(e[1]), 2), "backward_ms": round(e[1].elapsed_time(e[2]), 2),
            "cpu_port_forward_ms": round((t1 - t0) * 1e3, 1), "cpu_port_backward_ms": round((t2 - t1) * 1e3, 1), "cpu_cores": 1,
            "x_equals_cpu_port_bitwise": same,
            "kernel": "trisolve_band_forward_kernel / trisolve_band_backward_kernel (proper band factor: solver warps of 8 right-hand sides, "
                      "staging warps, no hand-over of a row between warps)",
            "note": "rows are sequential by construction: bound by the issue rate of the solver warps (forward) and by the chain of 32 "
                    "dependent additions the reference's order forces (backward)"}


# ---------------------------------------------------------------------------------------------
# CPU baseline = the reference's CPU path (C restatement, 1 thread) on a bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_reference_sample(kind, prm, n, dtype, budget_s, b_host_cols=None, fixed_rows=None):
    """Time the faithful restatement of Csr::mul_dense on a contiguous row sample of the workload.
    Returns (gflops, seconds, sample_rows, sample_nnz, description)."""
    from basic_sparse_matrix_b200 import gen
    from oracle.ref_cpu import OracleCsr
    if kind not in ("laplace3d", "laplace2d"):
        raise ValueError("cpu sample implemented for the Laplacian workloads")
    g = prm["g"]
    gz = g if kind == "laplace3d" else 1
    rows_total = g * g * gz
    reach = g * g if kind == "laplace3d" else g
    start = (rows_total // 2) // reach * reach          # interior block, aligned to a plane / grid line

    def run(nrows):
        r0, r1 = start, min(rows_total, start + nrows)
        v, ci, ri, _ = gen.laplacian(g, g, gz, r0, r1, dtype)
        cmin, cmax = int(ci.min()), int(ci.max()) + 1
        if b_host_cols is not None:
            cols = [c[cmin:cmax] for c in b_host_cols]
        else:
            blk = gen.dense_rows(rows_total, n, 5, gen.MODE_EXACT, 0.0, dtype, row_ids=np.arange(cmin, cmax))
            cols = [np.ascontiguousarray(blk[:, c]) for c in range(n)]
        # same arithmetic with rebased column indices; dims.cols stays the FULL k so the per-row
        # Vec::with_capacity(dims.cols) of get_row_compact (sparse.rs:254) costs what it costs
        a = OracleCsr.from_raw((r1 - r0, rows_total), v, ci - np.uint64(cmin), ri)
        t, _ = a.time_mul_dense_rows(cols, 0, r1 - r0, faithful=True, rhs_row_count=rows_total)
        return t, r1 - r0, int(ri[-1])

    if fixed_rows is not None:
        t, nr, nnz = run(fixed_rows)
    else:
        t, nr, nnz = run(4096)                           # calibrate
        want = int(min(rows_total - start, max(4096, 4096 * budget_s / max(t, 1e-6))))
        want = min(want, 1 << 21)
        if want > 8192:
            t, nr, nnz = run(want)
    gf = 2.0 * nnz * n / t / 1e9
    desc = (f"rows [{start},{start + nr}) of the workload ({nr} rows, {nnz} nnz) x {n} cols, faithful C restatement of "
            f"src/sparse.rs:426-446 incl. per-row Vec::with_capacity(cols) and the zero-dropping result Csr, 1 thread")
    return gf, t, nr, nnz, desc


def cpu_lean_parallel_sample(kind, prm, n, dtype, nrows, b_host_cols=None):
    """NOT the reference — context only: the same contraction in the same order on ALL host threads, dense
    output, no per-row allocation (oracle/ref_cpu.c ocsr_time_lean_parallel_*), on a row sample."""
    from basic_sparse_matrix_b200 import gen
    from oracle.ref_cpu import OracleCsr
    g = prm["g"]
    gz = g if kind == "laplace3d" else 1
    rows_total = g * g * gz
    reach = g * g if kind == "laplace3d" else g
    start = (rows_total // 2) // reach * reach
    r0, r1 = start, min(rows_total, start + nrows)
    v, ci, ri, _ = gen.laplacian(g, g, gz, r0, r1, dtype)
    cmin, cmax = int(ci.min()), int(ci.max()) + 1
    if b_host_cols is not None:
        cols = [c[cmin:cmax] for c in b_host_cols]
    else:
        blk = gen.dense_rows(rows_total, n, 5, gen.MODE_EXACT, 0.0, dtype, row_ids=np.arange(cmin, cmax))
        cols = [np.ascontiguousarray(blk[:, c]) for c in range(n)]
    a = OracleCsr.from_raw((r1 - r0, cmax - cmin), v, ci - np.uint64(cmin), ri)
    t, threads, _ = a.time_lean_parallel(cols, 0, r1 - r0)
    nnz = int(ri[-1])
    return {"value": round(2.0 * nnz * n / t / 1e9, 3), "unit": "GFLOP/s", "cores": threads, "seconds": round(t, 3),
            "sample_rows": r1 - r0, "note": "NOT the reference: lean multi-threaded variant of the same sum (dense output, "
                                            "no per-row allocation), for context"}


def main_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/ C port; the
    Rust original cannot be built here), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, prm, n, dt = WORKLOADS[args.workload]
    dtype = NP_DTYPE[dt]
    budget = 150.0 / max(1, args.steps + args.warmup)
    # size one step's sample from a calibration run, then time W + K steps of that sample
    _, t_cal, _, _, _ = cpu_reference_sample(kind, prm, n, dtype, 0.0, fixed_rows=2048)
    rows_per_step = int(max(1024, min(1 << 20, 2048 * budget / max(t_cal, 1e-6))))
    times, nnz_s, desc = [], 0, ""
    for i in range(args.warmup + args.steps):
        gf, t, nr, nnz_s, desc = cpu_reference_sample(kind, prm, n, dtype, 0.0, fixed_rows=rows_per_step)
        if i >= args.warmup:
            times.append(t)
    t = sum(times) / len(times)
    value = 2.0 * nnz_s * n / t / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": dt, "data": "synthetic",
            "config": {"workload": args.workload, "note": "each step = one bounded row sample of the workload"},
            "cpu_baseline": {"value": round(value, 5), "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": desc},
            "e2e": {"value": round(value, 5), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# the workloads reported under other_workloads at N=1 (the rest of WORKLOADS is for tools/sweep.py)
EXTRAS = ["laplace2d_2048_n1_f64", "rmat20_n64_f64", "rmat20_n64_f32", "band_1m_hb32_n32_f32", "band_1m_hb32_n1_f32",
          "laplace3d_256_n16_f64", "laplace3d_256_n8_f64", "laplace3d_256_n4_f64", "laplace3d_256_n1_f64",
          "laplace3d_252_n128_f64", "laplace2d_4096_n64_f64"]


def pinned(torch, shape, np_dtype):
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32, np.dtype(np.uint64): torch.int64}[np.dtype(np_dtype)]
    return torch.empty(shape, dtype=tdt).pin_memory().numpy().view(np_dtype)


def pcie_probe(torch, barrier, min_over_ranks, mb=512, reps=3):
    """What the platform gives every rank when ALL ranks copy at once: plain pinned-buffer copies (no library code), host->device
    only, device->host only, and both directions together. The ceiling the e2e numbers are to be read against."""
    n = mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for name in ("h2d", "d2h", "duplex"):
        torch.cuda.synchronize()
        if barrier:
            barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if name in ("h2d", "duplex"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if name in ("d2h", "duplex"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        dt = time.perf_counter() - t0
        out[name + "_gbs_per_direction_slowest_rank"] = round(min_over_ranks(n * reps / dt / 1e9), 1)
    out["note"] = f"all ranks at once, {mb} MB pinned buffers, torch copy_ on two streams; per direction, slowest rank"
    return out


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the gathered result (NCCL all-gather of C row blocks and the fused scatter) at N > 1")
    ap.add_argument("--gather", action="store_true", help="(default at N > 1; kept for compatibility)")
    ap.add_argument("--no-target", action="store_true", help="skip the north_star target case (x64) in the same run")
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--tune", default="", help="k=v,k=v overrides of bsm_tuning (sweeps)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return main_reference(args)

    import torch
    import torch.distributed as dist
    from basic_sparse_matrix_b200 import Csr, Dense, gen, gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:   # the literal call expands the result's row masks with host threads: N ranks share the host's cores
        os.environ.setdefault("BSM_PIPE_EXPAND_THREADS", str(max(1, min(8, (os.cpu_count() or 8) // (2 * world)))))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)      # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    barrier = (lambda: dist.barrier()) if world > 1 else None
    gpu.init(local)
    # the library enqueues on the stream it is given; torch.cuda.Event only sees torch's CURRENT
    # stream, so the whole bench runs inside one dedicated non-default torch stream
    stream = torch.cuda.Stream()
    gpu.set_stream(stream.cuda_stream)
    torch.cuda.set_stream(stream)
    peak, peak_src = load_peaks()

    def all_true(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    kind, prm, n, dt = WORKLOADS[args.workload]
    dtype = NP_DTYPE[dt]
    s = np.dtype(dtype).itemsize
    tuning = None
    if args.tune or args.algo != "auto":
        kw = {k: int(v, 0) for k, v in (kv.split("=") for kv in args.tune.split(",") if kv)}
        tuning = gpu.make_tuning(args.algo, **kw)

    # ---- operands: nnz-balanced row block of A on this rank, B replicated ---------------------
    ri_full = host_row_index(kind, prm)
    a_host = None
    if ri_full is not None:
        rows_total, nnz_total = len(ri_full) - 1, int(ri_full[-1])
        bounds = gpu.partition_rows(ri_full, world).astype(np.int64)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        A = make_device_csr(gpu, kind, prm, dtype, r0, r1)
    else:   # R-MAT: generated whole on the device; single-GPU only
        if world > 1:
            raise SystemExit("the R-MAT workload is generated on one device; run it with --gpus 1")
        A = make_device_csr(gpu, kind, prm, dtype)
        rows_total, nnz_total = A.info()["rows"], A.info()["nnz"]
        bounds = np.array([0, rows_total], np.int64)
        r0, r1 = 0, rows_total
        a_host = A.to_host().raw_parts()
    ai = A.info()
    reach = {"laplace3d": lambda: prm["g"] ** 2, "laplace2d": lambda: prm["g"], "band": lambda: prm["hb"]}.get(kind, lambda: 0)()
    B = gpu.DeviceDense.generate(ai["cols"], n, seed=B_SEED, mode=gen.MODE_EXACT, dtype=dtype)
    C = gpu.DeviceDense.alloc(ai["rows"], n, dtype)

    # ---- device-resident timing --------------------------------------------------------------------
    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        A.mul_dense(B, out=C, tuning=tuning)
    torch.cuda.synchronize()
    launches_w = gpu.kernel_launch_count()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        A.mul_dense(B, out=C, tuning=tuning)
        evs[i + 1].record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    clocks = sampler.stop()
    launches = gpu.kernel_launch_count() - launches_w
    info = gpu.last_launch_info()
    total_ms = evs[0].elapsed_time(evs[-1])
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms_max = max_over_ranks(total_ms)
    t_step = total_ms_max / args.steps * 1e-3
    value = 2.0 * nnz_total * n / t_step / 1e9

    # ---- parity of THIS run's result, on every rank: sampled rows of the rank's own C block against the CPU oracle ----
    ids = parity_row_ids(r0, r1, 64, reach)
    got = sample_dense_rows(gpu, C, ids - r0)
    ok_local = bitwise_equal(got, oracle_rows(kind, prm, n, dtype, ids, a_host))
    parity = {"rows_checked": int(len(ids)) * world, "bitwise": all_true(ok_local), "checker": "CPU oracle (oracle/ref_numpy.py: sequential sum in "
              "stored order, mul then add — src/sparse.rs:431-444) on rows of every rank's device-resident C block: first / last rows, "
              "rows one stencil reach inside the block, hashed rows", "ranks": world}

    # roofline of the dominant kernel on this rank (per launch)
    kref = k_ref_of(kind, prm, rows_total, r0, r1)
    bm_rank = bytes_min(ai["rows"], ai["nnz"], kref, n, s)
    t_launch = total_ms / args.steps * 1e-3 / max(1, info["passes"])
    achieved = bm_rank / max(1, info["passes"]) / t_launch / 1e9 if info["passes"] else 0.0
    # DRAM traffic of the kernel from the committed ncu capture — only if it is a capture of THIS launch configuration
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            ent = json.load(open(tp)).get(args.workload)
            sig = {k: info[k] for k in ("algo", "lanes_per_row", "reg_tiles", "reg_flavour", "rows_per_warp")}
            if isinstance(ent, dict) and ent.get("launch") == sig:
                traffic, traffic_src = ent.get("dram_bytes"), ent.get("source")
        except Exception:
            traffic = None
    bm_total = bytes_min(rows_total, nnz_total, rows_total, n, s)
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "kernel": KERNEL_NAME.get(info["algo"], "?"),
                "algorithmic_bytes_per_launch": int(bm_rank / max(1, info["passes"])),
                "launch_ms_avg": round(t_launch * 1e3, 4), "launch_ms_best": round(min(per) / max(1, info["passes"]), 4),
                "peak_source": peak_src}

    # ---- the gathered result (the only collective of the path; optional for callers, always measured here at N > 1) ----
    gather = None
    if world > 1 and not args.no_gather:
        gather = {}
        try:
            uid = [gpu.Comm.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            comm = gpu.Comm.init(uid[0], world, rank)
            full = gpu.DeviceDense.alloc(rows_total, n, dtype)
            comm.allgather_rows(C, bounds.astype(np.uint64), full)       # warm-up (NCCL channel setup)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            comm.allgather_rows(C, bounds.astype(np.uint64), full)
            e1.record()
            torch.cuda.synchronize()
            g_ms = max_over_ranks(e0.elapsed_time(e1))
            # rows of the GATHERED buffer against the oracle: a few rows of EVERY rank's block, not only this rank's
            gids = np.unique(np.concatenate([parity_row_ids(int(bounds[q]), int(bounds[q + 1]), 12, reach) for q in range(world)]))
            want_g = oracle_rows(kind, prm, n, dtype, gids, a_host)
            ok_nccl = all_true(bitwise_equal(sample_dense_rows(gpu, full, gids), want_g))
            gather.update({"allgather_ms": round(g_ms, 3), "bytes_received_per_gpu": int((rows_total - ai["rows"]) * n * s),
                           "spmm_plus_allgather_ms": round(total_ms_max / args.steps + g_ms, 3),
                           "allgather_parity": {"rows_checked": int(len(gids)), "bitwise": ok_nccl,
                                                "checker": "CPU oracle on rows of every rank's block in every rank's gathered buffer"}})
            # the same gathered product WITHOUT the collective: the SpMM kernel stores every C row into every
            # rank's full buffer (peer memory over NVLink, mapped with CUDA IPC): multiply + gather in one kernel
            try:
                fi = full.info()
                handles = [None] * world
                dist.all_gather_object(handles, full.ipc_export())
                peers = [gpu.DeviceDense.ipc_open(handles[q], rows_total, n, fi["ld"], dtype) for q in range(world) if q != rank]
                dests = [full] + peers
                torch.cuda.synchronize()
                dist.barrier()
                gpu.fill_zero(full)                                    # so that the check below sees only what the fused kernel wrote
                torch.cuda.synchronize()
                dist.barrier()
                A.mul_dense_scatter(B, dests, r0)                      # warm-up (maps peer pages)
                comm.barrier()
                torch.cuda.synchronize()
                dist.barrier()
                reps = 3
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(reps):
                    A.mul_dense_scatter(B, dests, r0)
                    comm.barrier()
                f1.record()
                torch.cuda.synchronize()
                f_ms = max_over_ranks(f0.elapsed_time(f1) / reps)
                ok_fused = all_true(bitwise_equal(sample_dense_rows(gpu, full, gids), want_g))
                gather.update({"fused_scatter_ms": round(f_ms, 3),
                               "fused_parity": {"rows_checked": int(len(gids)), "bitwise": ok_fused,
                                                "checker": "CPU oracle; the buffer was zeroed before the fused kernel ran"},
                               "fused_path": "bsm_spmm_scatter: P2P stores of every C row to all ranks' full buffers (CUDA IPC), then a 4-byte NCCL barrier"})
                dist.barrier()
                for h in peers:
                    h.close()
            except Exception as ex:
                gather["fused_scatter_error"] = str(ex)[:300]
            full.close()
            comm.close()
        except Exception as ex:
            gather["error"] = str(ex)[:300]

    # ---- north_star target case in the same run: the same matrix x 64 columns, f64 (>= 100 M nnz x 64) ------------------
    target = None
    if not args.no_target and kind == "laplace3d" and dt == "f64" and n != 64:
        try:
            nt = 64
            Bt = gpu.DeviceDense.generate(ai["cols"], nt, seed=B_SEED, mode=gen.MODE_EXACT, dtype=dtype)
            Ct = gpu.DeviceDense.alloc(ai["rows"], nt, dtype)
            t_ms, t_per = time_device_steps(torch, A, Bt, Ct, max(5, args.steps // 2), 3, None, barrier)
            t_info = gpu.last_launch_info()
            steps_t = max(5, args.steps // 2)
            t_ms_max = max_over_ranks(t_ms) / steps_t
            ok_t = all_true(bitwise_equal(sample_dense_rows(gpu, Ct, ids - r0), oracle_rows(kind, prm, nt, dtype, ids, a_host)))
            bm_t = bytes_min(rows_total, nnz_total, rows_total, nt, s)
            bm_t_rank = bytes_min(ai["rows"], ai["nnz"], kref, nt, s)
            target = {"workload": f"laplace3d_{prm['g']}_n64_f64", "ncols": nt, "nnz": nnz_total, "ms_per_step": round(t_ms_max, 4),
                      "gflops": round(2.0 * nnz_total * nt / (t_ms_max * 1e-3) / 1e9, 1),
                      "roofline_frac": round(bm_t_rank / (t_ms / steps_t * 1e-3) / 1e9 / peak, 4),
                      "roofline_frac_job": round(bm_t / (t_ms_max * 1e-3) / 1e9 / (peak * world), 4),
                      "north_star_bar": ">= 0.60 of the HBM roofline on 1 GPU, >= 6x at 8 GPUs",
                      "parity_bitwise": ok_t, "launch": {k: t_info[k] for k in ("algo", "lanes_per_row", "reg_tiles", "reg_flavour", "grid", "block")}}
            Ct.close()
            if world > 1:
                # the same run's ONE-GPU time of the whole problem (rank 0, the others wait), for the speed-up
                one_ms = 0.0
                if rank == 0:
                    A1 = make_device_csr(gpu, kind, prm, dtype)
                    C1 = gpu.DeviceDense.alloc(rows_total, nt, dtype)
                    o_ms, _ = time_device_steps(torch, A1, Bt, C1, steps_t, 3)
                    one_ms = o_ms / steps_t
                    A1.close()
                    C1.close()
                dist.barrier()
                one_ms = max_over_ranks(one_ms)
                target.update({"one_gpu_ms_same_run": round(one_ms, 4), "speedup_vs_one_gpu_same_run": round(one_ms / t_ms_max, 3)})
            Bt.close()
        except Exception as ex:
            target = {"error": str(ex)[:300]}

    # ---- end to end through the reference-facing C-ABI calls, from pinned HOST buffers --------------------
    e2e = e2e_dense = None
    host_cols = None
    if not args.no_e2e:
        torch_dt = torch.float64 if dt == "f64" else torch.float32
        # host operands in the REFERENCE layout: Csr fields (usize indices) and Dense columns
        hA = A.to_host()                                    # rows [r0,r1) as a finalised Csr
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        hv, hci, hri = (pin(x) for x in hA.raw_parts())
        hA = Csr.from_raw_parts((ai["rows"], ai["cols"]), hv, hci, hri)
        host_cols = [torch.empty(ai["cols"], dtype=torch_dt).pin_memory().numpy() for _ in range(n)]
        B.to_host(Dense.from_columns_nocopy(host_cols))     # the host copy of B (outside the timed region)
        hB = Dense.from_columns_nocopy(host_cols)
        # the calls upload only the window of B rows this rank's A block references (all of B at N=1)
        h2d = hv.nbytes + hci.nbytes + hri.nbytes + kref * n * s
        chk_ids = parity_row_ids(r0, r1, 64, reach)
        want_rows = oracle_rows(kind, prm, n, dtype, chk_ids, a_host)

        def timed(fn, steps):
            gpu.phase_timers(enable=True)                   # reset
            fn()                                            # warm-up (stream-ordered pool, pinned pages)
            torch.cuda.synchronize()
            gpu.phase_timers()                              # discard the warm-up's phases
            if barrier:
                barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            torch.cuda.synchronize()
            te = (time.perf_counter() - t0) / steps
            ph = {k: round(v / steps * 1e3, 2) for k, v in gpu.phase_timers(enable=False).items()}
            return max_over_ranks(te), ph

        steps_e = max(1, args.e2e_steps)
        try:
            pcie = pcie_probe(torch, barrier, min_over_ranks)
        except Exception as ex:
            pcie = {"error": str(ex)[:200]}
        # (1) dense result
        out_cols = [torch.empty(ai["rows"], dtype=torch_dt).pin_memory().numpy() for _ in range(n)]
        hC = Dense.from_columns_nocopy(out_cols)
        te, ph = timed(lambda: hA.mul_dense_into(hB, hC, algo=args.algo), steps_e)
        got_rows = np.stack([np.array([c[r - r0] for c in out_cols]) for r in chk_ids]).astype(dtype)
        e2e_dense = {"value": round(2.0 * nnz_total * n / te / 1e9, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d),
                     "d2h_bytes_per_step": int(sum(c.nbytes for c in out_cols)), "ms_per_step": round(te * 1e3, 2), "steps": steps_e,
                     "parity": {"rows_checked": int(len(chk_ids)) * world, "bitwise": all_true(bitwise_equal(got_rows, want_rows)), "checker": "CPU oracle"},
                     "phases_ms_rank0": ph,
                     "path": "Csr.mul_dense_into -> bsm_mul_dense_host_dense_*: host Csr + host Dense columns in, host Dense columns out (no zero-drop)"}
        del out_cols, hC
        # (2) the literal call: zero-dropped Csr result into pinned arrays sized for the worst case (every output non-zero)
        cap = ai["rows"] * n
        need_gb = cap * (s + 8) / 1e9
        avail = mem_available_gb()
        if avail is not None and need_gb > 0.8 * avail:
            e2e = {"skipped": f"result arrays need {need_gb:.1f} GB of pinned host memory, {avail:.1f} GB available"}
        else:
            ov, oc, orow = pinned(torch, cap, dtype), pinned(torch, cap, np.uint64), pinned(torch, ai["rows"] + 1, np.uint64)
            res = [None]

            def step_csr():
                res[0] = hA.mul_dense_csr_into(hB, ov, oc, orow, algo=args.algo)

            te, ph = timed(step_csr, steps_e)
            rv, rc, rr = res[0].raw_parts()
            got_rows = csr_rows_dense(rv, rc, rr, chk_ids, r0, n, dtype)
            nnz_out = int(rr[-1])
            # what actually crosses PCIe: values + row_index, and the columns either as usize values (8 B per entry) or — the default —
            # as one keep-bit per output, expanded into col_index by host threads inside the call (csrc/pipeline.cu, ColumnExpand)
            expand_threads = os.environ.get("BSM_PIPE_EXPAND_THREADS", "")
            by_masks = expand_threads != "0"
            col_bytes = ai["rows"] * ((n + 63) // 64) * 8 if by_masks else nnz_out * 8
            e2e = {"value": round(2.0 * nnz_total * n / te / 1e9, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(nnz_out * s + col_bytes + (ai["rows"] + 1) * 8), "ms_per_step": round(te * 1e3, 2), "steps": steps_e,
                   "result_bytes_in_host_memory": int(nnz_out * (s + 8) + (ai["rows"] + 1) * 8),
                   "columns_on_the_wire": ("row masks (1 bit per output), expanded by host threads inside the call"
                                           + (f" (BSM_PIPE_EXPAND_THREADS={expand_threads})" if expand_threads else " (default: 8 threads)"))
                   if by_masks else "usize values",
                   "result": "Csr (zero-dropped, usize indices) — the reference's return type", "result_nnz_rank0": nnz_out,
                   "parity": {"rows_checked": int(len(chk_ids)) * world, "bitwise": all_true(bitwise_equal(got_rows, want_rows)),
                              "checker": "CPU oracle; rows of the returned Csr densified (dropped zeros = 0)"},
                   "phases_ms_rank0": ph, "numa": numa, "pcie_probe": pcie,
                   "path": "Csr.mul_dense_csr_into -> bsm_mul_dense_host_into_*: the literal Csr::mul_dense (sparse.rs:426-446): host Csr + host Dense "
                           "columns in, zero-dropped host Csr out; pipeline = A up | B up in row chunks + transpose | per row block SpMM + count/scan/scatter | "
                           "values + row masks + row_index down, col_index written by host threads from the masks; pinned host buffers"}
            del ov, oc, orow, res

    # ---- CPU baseline (rank 0, N=1 only): the reference's CPU path on a bounded sample ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and kind in ("laplace3d", "laplace2d"):
        gf, t, nr, nnz_s, desc = cpu_reference_sample(kind, prm, n, dtype, 12.0, b_host_cols=host_cols)
        cpu = {"value": round(gf, 5), "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": desc,
               "seconds": round(t, 2), "host_cores_available": os.cpu_count()}
        try:
            cpu["lean_all_cores"] = cpu_lean_parallel_sample(kind, prm, n, dtype, 1 << 20, b_host_cols=host_cols)
        except Exception as ex:
            cpu["lean_all_cores"] = {"error": str(ex)[:200]}

    # ---- free the headline operands, then the secondary workloads (N=1 only) ----------------------------
    for h in (A, B, C):
        h.close()
    host_cols = None
    extras = []
    if rank == 0 and world == 1 and not args.no_extras:
        for name in EXTRAS:
            if name == args.workload:
                continue
            try:
                extras.append(run_extra(torch, gpu, name, 10, 3, peak))
            except Exception as ex:   # keep the headline line even if a side workload fails
                extras.append({"workload": name, "error": str(ex)[:200]})
        try:   # config 5's product with the OPT-IN fused arithmetic (tolerance-level agreement), next to the bit-exact default above
            from basic_sparse_matrix_b200 import _lib
            fused = gpu.make_tuning("auto", flags=_lib.TUNE_A_EVICT_FIRST | _lib.TUNE_C_STREAMING | _lib.TUNE_FUSED)
            extras.append(run_extra(torch, gpu, "band_1m_hb32_n32_f32", 10, 3, peak, fused, "band_1m_hb32_n32_f32 [BSM_TUNE_FUSED opt-in]"))
        except Exception as ex:
            extras.append({"workload": "band_1m_hb32_n32_f32 [BSM_TUNE_FUSED opt-in]", "error": str(ex)[:200]})
        try:
            extras.append(run_solve(torch, gpu))
        except Exception as ex:
            extras.append({"workload": "band solve", "error": str(ex)[:200]})
        try:
            extras.append(run_config1(torch, gpu, peak))
        except Exception as ex:
            extras.append({"workload": "bench_as_written", "error": str(ex)[:200]})

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(total_ms_max / args.steps, 4), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": dt, "data": "synthetic",
                "config": {"workload": args.workload, "rows": rows_total, "nnz": nnz_total, "ncols": n,
                           "partition": f"nnz-balanced row blocks x{world}, B replicated, no data-path collective",
                           "l2": "operands >> L2 (126 MB), no flush needed", "values": "exact dyadic (k/1024), hash-generated on device",
                           "algo": ALGO_NAME.get(info["algo"], "?"), "launch": info},
                "parity": parity, "north_star_target": target,
                "effective_gbs": round(bm_total / t_step / 1e9, 1),
                "roofline_frac_job": round(bm_total / t_step / 1e9 / (peak * world), 4),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_dense": e2e_dense, "gpu_launches": int(launches), "clocks": clocks,
                "ms_per_step_best": round(min(per), 4), "gather": gather, "other_workloads": extras}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
