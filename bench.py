#!/usr/bin/env python
"""bench.py -- collocation Jacobian nnz/s of the Radau transcription hot path on B200.

Workload (BASELINE.json config 4): batched quadrotor MPC, 4096 independent OCP instances per
GPU on a shared 8x8 LGR mesh (ns=12, nc=4, N=64; n=1038, m=769, nnz_jac=19970 per instance),
instance i seeded with PCG64(5+i).  A "step" is one fused eval_g + eval_jac_g(values) pass over
the whole batch (the reference's GetAllCons + GetConsJacbi, LpNLPWrapper.cpp:34,:230) by the
reference's own finite-difference scheme.  Instances shard over ranks with no data-path
collective (weak scaling: 4096 instances per GPU).

  value      nnz_jac x instances x steps / device time, inputs resident in HBM
  e2e        same through the C-ABI host call (pinned host buffers, H2D of x and D2H of g and
             values inside the timed region)
  roofline   k_cons_jac (dominant kernel): algorithmic bytes / CUDA-event duration vs the
             measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference: the reference's OWN transcription sources on the host cores of
             the same box (kind "reference": oracle/_ref/liblpopc_ref.so, the unmodified translation
             units of /root/reference compiled against this repository's Armadillo stand-in
             oracle/ref_shim/ at -O2 -- real Armadillo/BLAS and IPOPT are not in the image; one forked
             single-threaded worker per core because the reference is not re-entrant; kind "port":
             the oracle/ restatement, only when no oracle/_ref build exists), same workload.
  strong     the same step on BASELINE config 4 as written: 4096 instances in TOTAL, sharded over
             the N ranks (lpopc_b200.batch.shard_range), device-timed, max over ranks.
  c5, hessian, hessian_probed: side measurements on rank 0 with their own roofline objects.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

INSTANCES_PER_GPU = 4096
SOLVE_INSTANCES = 4096  # per GPU, for the solves/s side measurement
MESH_INTERVALS, MESH_NODES = 8, 8
NBUF = 4  # rotating input sets so that x is not served from L2 between steps
# dram__bytes_read.sum + dram__bytes_write.sum of one k_cons_jac launch of this workload (4096 instances)
# from the committed `ncu --set full` capture (35.23 MB read + 621.77 MB written; below the 713 MB algorithmic figure
# because part of the last stores is still in L2 at kernel end).  DRAM counters cannot be read outside a profiler, so
# this is NOT measured in the bench run; the line says where it comes from.
NCU_TRAFFIC_BYTES = 657005824
NCU_TRAFFIC_SOURCE = "profiles/r02_cons_jac_full.txt (ncu --set full of this kernel on this workload; not measured in this run)"
# FP64 issue rate of one B200 SM, measured with scripts/dev/fp64_peak.cu (independent DFMA chains, 32 warps per SM):
# 1.81 warp instructions per cycle per SM = 33.7 TFLOP/s over 148 SMs at 1965 MHz.  The side kernels below are bound by
# FP64 issue and latency, not by HBM, so their roofline objects carry this second bound; the executed FP64 warp
# instructions (DADD + DMUL + DFMA + DSETP) come from the committed ncu captures, not from this run.
FP64_WARP_INST_PER_CLK_PER_SM = 1.81
NCU_FP64_WARP_INSTS = {"hessian": (82706437, "profiles/r02_hess_tiled_full.txt (k_hess_tiled, 4096 quadrotor instances)"),
                       "c5": (63393752, "profiles/r02_cons_jac_rows_full.txt (k_cons_jac_rows, config 5)")}
STRONG_TOTAL = 4096  # BASELINE config 4 as written: 4096 instances in total, sharded over the ranks
C5_INTERVALS, C5_NODES = 10000, 10  # BASELINE config 5: synthetic ns=20 / nc=6 dynamics on 100k LGR nodes


def quadrotor_problem():
    from lpopc_b200 import examples
    return examples.quadrotor(intervals=MESH_INTERVALS, nodes=MESH_NODES)


def make_inputs(op, lgr_points, first, count):
    """x[count, n]: per-instance MPC guess (line from the perturbed initial state to the target)
    plus a small perturbation, PCG64(5 + instance)."""
    base = op.guess(lgr_points)
    n = base.size
    ph = op.phases[0]
    N = ph.GetTotalNodes()
    ns = len(ph.statemin)
    tau = np.concatenate([np.asarray(lgr_points[0]), [1.0]])
    ramp = 0.5 * (1.0 - tau)  # weight of the initial state along the guess line
    X = np.empty((count, n))
    for i in range(count):
        rng = np.random.Generator(np.random.PCG64(5 + first + i))
        dx0 = 0.2 * rng.uniform(-1, 1, ns)
        x = base.copy()
        for j in range(ns):
            x[j * (N + 1):(j + 1) * (N + 1)] += dx0[j] * ramp
        x[: n - 2] += 0.05 * rng.uniform(-1, 1, n - 2)
        X[i] = x
    return X


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML back to back
    (same counters as the nvidia-smi clocks line of B200_PROFILING.md; nvidia-smi's own 100 ms
    period is longer than a short timed region)."""

    def __init__(self, index, period_s=0.0):
        self.index = index
        self.period_s = period_s  # 0: poll back to back (millisecond-long timed regions); the polling thread holds the GIL
                                  # while it runs, so regions that execute Python for seconds use a few milliseconds
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception as e:  # NVML missing: report it, do not fake clocks
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, rs))
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(self.period_s)  # 0 = yield only; NVML itself takes a few tens of microseconds per query

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        nv = self.nv
        smax = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}
        reasons = set()
        for _, rs in self.samples:
            for nm, bit in bits.items():
                if rs & bit:
                    reasons.add(nm)
        sm = [x for x, _ in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(smax), "reasons": sorted(reasons),
                "samples": len(sm), "source": "NVML polled %s inside the timed region" % ("back to back" if self.period_s == 0 else "every %g ms" % (1e3 * self.period_s))}


REF_LIB = os.path.join(ROOT, "oracle", "_ref", "liblpopc_ref.so")
_worker = {}


def _ref_worker_init():
    from oracle_lib import RefOracle
    _worker["o"] = RefOracle(quadrotor_problem())


def _ref_worker_eval(span):
    lo, hi = span
    g, v = _worker["o"].eval_g_jac_batch(_worker["X"][lo:hi], nthreads=1)
    return float(v[:, 0].sum())


class CpuPath:
    """The reference's CPU implementation of the path on `threads` host cores.

    kind "reference": the reference's OWN sources (oracle/_ref/liblpopc_ref.so, built from
    /root/reference against oracle/ref_shim/ by oracle/ref_build.mk).  The reference is
    single-threaded and not re-entrant (function-static caches), so the cores are used by forked
    worker processes, each evaluating a contiguous share of the instances.
    kind "port": the restatement (oracle/) with one std::thread per core, when no _ref was built."""

    def __init__(self, op, X, threads):
        self.X, self.threads = X, threads
        from oracle_lib import Oracle
        self.nnz = Oracle(op).nnz_jac
        self.kind = "reference" if os.path.exists(REF_LIB) else "port"
        self.lib = None
        if self.kind == "reference":
            import hashlib
            import multiprocessing as mp
            from oracle_lib import ref_lib
            ref_lib()  # dlopen in THIS process too (the forked workers inherit the mapping): the library that is timed
            self.lib = {"path": os.path.relpath(REF_LIB, ROOT), "sha256": hashlib.sha256(open(REF_LIB, "rb").read()).hexdigest(),
                        "built_from": "/root/reference/Lpopc/src (unmodified translation units) + oracle/ref_shim (Armadillo stand-in), g++ -O2, oracle/ref_build.mk"}
            _worker["X"] = X  # inherited by the forked workers
            self.pool = mp.get_context("fork").Pool(threads, initializer=_ref_worker_init)
            per = (len(X) + threads - 1) // threads
            self.spans = [(i, min(len(X), i + per)) for i in range(0, len(X), per)]
        else:
            self.o = Oracle(op)

    def step(self):
        if self.kind == "reference":
            self.pool.map(_ref_worker_eval, self.spans, chunksize=1)
        else:
            self.o.eval_g_jac_batch(self.X, nthreads=self.threads)

    def close(self):
        if self.kind == "reference":
            self.pool.close()
            self.pool.join()

    def baseline(self, value, threads, sample):
        d = {"value": value, "unit": "nnz/s", "cores": threads, "kind": self.kind, "sample": sample}
        if self.lib:
            d["library"] = self.lib
        return d

    def describe(self):
        if self.kind == "reference":
            return ("the reference's own sources (Core/LpNLPWrapper.cpp, LpFiniteDifferenceDerive.cpp, ... compiled against the "
                    "Armadillo stand-in oracle/ref_shim), %d forked single-threaded workers" % self.threads)
        return "oracle/ restatement of the reference path, %d host threads (no oracle/_ref build available)" % self.threads


def cpu_reference_rate(op, X, threads, budget_s=10.0):
    """(nnz/s, seconds, passes, CpuPath) over the whole batch X, repeated for about budget_s seconds."""
    cp = CpuPath(op, X, threads)
    cp.step()  # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        cp.step()
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    cp.close()
    return cp.nnz * len(X) * reps / dt, dt, reps, cp


def run_reference(args, rank, world):
    if rank != 0:
        return
    op = quadrotor_problem()
    from oracle_lib import Oracle
    o = Oracle(op)
    pts = [o.tables(0)["points"]]
    threads = os.cpu_count() or 1
    sample = INSTANCES_PER_GPU
    X = make_inputs(op, pts, 0, sample)
    cp = CpuPath(op, X, threads)
    for _ in range(max(1, args.warmup)):
        cp.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cp.step()
    dt = time.perf_counter() - t0
    cp.close()
    value = o.nnz_jac * sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "collocation Jacobian nnz/s", "value": value, "unit": "nnz/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(o, world),
        "cpu_baseline": cp.baseline(value, threads, "all %d quadrotor instances of one GPU's share per step, on the host; %s" % (sample, cp.describe())),
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def hbm_roofline(algorithmic_bytes, ms, fp64_key=None, sm_mhz=1965.0):
    """roofline object of a side measurement: algorithmic bytes per evaluation / device time against the HBM peak, and --
    for the compute-bound side kernels -- the FP64 issue floor of the kernel's executed FP64 instructions."""
    peak, src = hbm_peak()
    ach = algorithmic_bytes / (ms * 1e-3) / 1e9
    out = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "bytes": algorithmic_bytes, "peak_source": src,
           "traffic": None}
    if fp64_key in NCU_FP64_WARP_INSTS:
        insts, where = NCU_FP64_WARP_INSTS[fp64_key]
        floor_ms = insts / (FP64_WARP_INST_PER_CLK_PER_SM * 148 * sm_mhz * 1e6) * 1e3
        out["fp64_issue"] = {"warp_insts": insts, "warp_insts_source": where + "; not measured in this run",
                             "peak_warp_inst_per_clk_per_sm": FP64_WARP_INST_PER_CLK_PER_SM, "peak_source": "scripts/dev/fp64_peak.cu on this pool's B200",
                             "floor_ms": floor_ms, "frac": floor_ms / ms}
    return out


class _Sizes:
    def __init__(self, n, m, nnz_jac, nnz_h):
        self.n, self.m, self.nnz_jac, self.nnz_h = n, m, nnz_jac, nnz_h


def workload_config(o, world):
    cfg = {"workload": "batched quadrotor MPC (BASELINE config 4): %d OCP instances per GPU, shared %dx%d LGR mesh, "
                       "fused eval_g+eval_jac_g by forward differences" % (INSTANCES_PER_GPU, MESH_INTERVALS, MESH_NODES),
           "instances_per_gpu": INSTANCES_PER_GPU, "instances_total": INSTANCES_PER_GPU * world,
           "n": o.n, "m": o.m, "nnz_jac": o.nnz_jac, "nnz_h": o.nnz_h, "parallelism": "instances sharded x%d, no data-path collective" % world,
           "l2": "%d rotating input sets + %.0f MB written per step (> 126 MB L2)" % (NBUF, 8e-6 * INSTANCES_PER_GPU * (o.m + o.nnz_jac))}
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the Hessian / large-mesh side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--dense-return", action="store_true", help="e2e: return the whole Jacobian head over PCIe (sparse_return = 0)")
    ap.add_argument("--no-persistent", action="store_true", help="e2e: the host rewrites the constant tail and the zero fill on every call")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="e2e: pipeline depth of the host-pointer batch call (0 = the library's default)")
    ap.add_argument("--return-mode", type=int, default=-1, help="e2e: 0 = zero-copy stores over PCIe, 1 = one strided DMA copy per run of non-zero segments; -1 = the library's default")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly one JSON line: anything libraries print on fd 1 meanwhile (NCCL's version banner,
    # warnings) is routed to stderr until rank 0 prints the result
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    cpu = None
    if rank == 0 and not args.no_cpu:
        # CPU leg first: it forks worker processes, which must happen before CUDA is initialised
        from oracle_lib import Oracle as _O
        _op = quadrotor_problem()
        _X = make_inputs(_op, [_O(_op).tables(0)["points"]], 0, INSTANCES_PER_GPU)
        cpu = cpu_reference_rate(_op, _X, os.cpu_count() or 1)

    # bind this rank to the CPUs local to its GPU (NVML's ideal affinity) BEFORE any pinned buffer or host
    # thread exists: the host-pointer call is PCIe/host-memory bound and with 8 ranks on one box remote-NUMA
    # pinned buffers cost more than half of its bandwidth
    affinity = "unchanged"
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        affinity = "gpu-local (nvmlDeviceSetCpuAffinity), %d cpus" % len(os.sched_getaffinity(0))
    except Exception as e:  # containers without the permission: keep going
        affinity = "unchanged (%s)" % type(e).__name__

    import torch
    import torch.distributed as dist
    from lpopc_b200 import batch, nlp
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)

    op = quadrotor_problem()
    # setup broadcast: every rank transcribes on rank 0's mesh (a few hundred bytes over NCCL)
    mp, nd = batch.broadcast_mesh(op.phases[0].meshpoints, op.phases[0].nodesperinterval, dist if world > 1 else None, dev)
    op.phases[0].set_mesh(mp, nd)
    g = nlp.TranscribedNLP(op)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    # ranks share the host cores; 8 fill threads saturate the host's write bandwidth (scripts/dev/e2e_probe.py), more only
    # compete with the driver's own threads
    g.set_option("host_threads", max(1, min(8, len(os.sched_getaffinity(0)) // world)))
    n, m, nnz, nnz_h = g.get_nlp_info()
    nb = INSTANCES_PER_GPU
    first = rank * nb

    pts = g.lgr_points()  # composite LGR nodes of the shared mesh, from the product's own tables
    X = make_inputs(op, pts, first, nb)

    xs = [torch.from_numpy(X + 1e-3 * k).to(dev) for k in range(NBUF)]
    d_g = torch.empty((nb, m), dtype=torch.float64, device=dev)
    d_v = torch.empty((nb, nnz), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step(k):
        g.eval_g_jac_dev(nb, xs[k % NBUF].data_ptr(), d_g.data_ptr(), d_v.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # clock window: one NVML query takes milliseconds, the K timed steps a few, so the sampler also watches the same
    # step running back to back (untimed) right before the timed region, until it holds >= 60 samples
    window_steps = 0
    t_w = time.perf_counter()
    while len(sampler.samples) < 60 and time.perf_counter() - t_w < 3.0 and sampler.thread is not None:
        for k in range(50):
            step(k)
        torch.cuda.synchronize()
        window_steps += 50
    in_window = len(sampler.samples)
    g.set_option("time_kernels", 1)
    l0 = g.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = g.kernel_launches - l0
    kern_ms, kern_cnt = g.kernel_time("cons_jac")
    g.set_option("time_kernels", 0)
    clocks = sampler.stop()
    clocks["samples_in_timed_region"] = max(0, clocks.get("samples", 0) - in_window)
    clocks["window"] = "%d untimed steps of the same kernel back to back immediately before the %d timed steps, then the timed steps" % (window_steps, args.steps)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all = float(tmax.item())
    value = nnz * nb * world * args.steps / (ms_all * 1e-3)

    # ---- end to end through the C-ABI host entry point (pinned host buffers) ----
    hx = [torch.from_numpy(X + 1e-3 * k).pin_memory() for k in range(2)]
    hg = torch.empty((nb, m), dtype=torch.float64).pin_memory()
    hv = torch.empty((nb, nnz), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    if args.dense_return:
        g.set_option("sparse_return", 0)
    # the caller (like IPOPT's TNLPAdapter) hands the same values array to every call and leaves it alone in
    # between: constant tail and zero fill stay in place, the host threads write nothing after the second call
    g.set_option("persistent_values", 0 if args.no_persistent else 1)
    if args.return_mode >= 0:
        g.set_option("return_mode", args.return_mode)
    if args.e2e_chunks > 0:
        g.set_option("e2e_chunks", args.e2e_chunks)
    for k in range(3):
        g.eval_g_jac_batch_ptr(nb, hx[k % 2].data_ptr(), hg.data_ptr(), hv.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        g.eval_g_jac_batch_ptr(nb, hx[k % 2].data_ptr(), hg.data_ptr(), hv.data_ptr())
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nnz * nb * world * e2e_steps / float(te.item())
    sparse_calls, sparse_on, sparse_fixups = g.stat("sparse_calls"), g.stat("sparse_on_doubles"), g.stat("sparse_fixups")
    persistent_hits = g.stat("persistent_hits")
    # PCIe floor of that call: one D2H copy of the bytes it returns (pinned), H2D of x in the other direction alongside
    d2h_bytes = 8 * nb * (m + (sparse_on if sparse_calls else nnz))
    probe_src = torch.empty(d2h_bytes // 8, dtype=torch.float64, device=dev)
    probe_dst = torch.empty(d2h_bytes // 8, dtype=torch.float64).pin_memory()
    side = torch.cuda.Stream()
    probe_dst.copy_(probe_src, non_blocking=True)
    torch.cuda.synchronize()
    barrier()  # all ranks copy at the same time: the floor includes what the ranks take from each other on the host side
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for k in range(3):
        with torch.cuda.stream(side):
            xs[0].copy_(hx[k % 2], non_blocking=True)
        probe_dst.copy_(probe_src, non_blocking=True)
    pe1.record()
    torch.cuda.synchronize()
    pcie_ms_rank = pe0.elapsed_time(pe1) / 3
    pcie_t = torch.tensor([pcie_ms_rank], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pcie_t, op=dist.ReduceOp.MAX)
    pcie_ms = float(pcie_t.item())
    del probe_src, probe_dst
    # the host-pointer call must deliver exactly what the device-resident call computes
    chk = torch.from_numpy(X + 1e-3 * ((e2e_steps - 1) % 2)).to(dev)
    g.eval_g_jac_dev(nb, chk.data_ptr(), d_g.data_ptr(), d_v.data_ptr())
    torch.cuda.synchronize()
    if not (torch.equal(hv.view(torch.int64), d_v.cpu().view(torch.int64)) and torch.equal(hg.view(torch.int64), d_g.cpu().view(torch.int64))):
        raise RuntimeError("host-pointer and device-resident evaluations disagree")
    del chk

    # ---- result gather (objective per instance; 8 B per instance over NCCL) ----
    d_f = torch.empty(nb, dtype=torch.float64, device=dev)
    g.eval_f_dev(nb, xs[0].data_ptr(), d_f.data_ptr())
    torch.cuda.synchronize()
    all_f = batch.gather_results(d_f.cpu().numpy(), nb * world, dist if world > 1 else None, dev)

    # ---- BASELINE config 4 as written: STRONG_TOTAL instances in total, sharded over the ranks ----
    slo, shi = batch.shard_range(STRONG_TOTAL, rank, world)
    snb = shi - slo
    sX = make_inputs(op, pts, slo, snb) if world > 1 else X
    sxs = [torch.from_numpy(sX + 1e-3 * k).to(dev) for k in range(NBUF)] if world > 1 else xs
    for k in range(args.warmup):
        g.eval_g_jac_dev(snb, sxs[k % NBUF].data_ptr(), d_g.data_ptr(), d_v.data_ptr())
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for k in range(args.steps):
        g.eval_g_jac_dev(snb, sxs[k % NBUF].data_ptr(), d_g.data_ptr(), d_v.data_ptr())
    s1.record()
    barrier()
    tsm = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    # the same shard through the host-pointer call
    shx = [torch.from_numpy(sX + 1e-3 * k).pin_memory() for k in range(2)] if world > 1 else hx
    for k in range(3):
        g.eval_g_jac_batch_ptr(snb, shx[k % 2].data_ptr(), hg.data_ptr(), hv.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        g.eval_g_jac_batch_ptr(snb, shx[k % 2].data_ptr(), hg.data_ptr(), hv.data_ptr())
    barrier()
    tse = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tsm, op=dist.ReduceOp.MAX)
        dist.all_reduce(tse, op=dist.ReduceOp.MAX)
    strong = {"scaling": "strong", "instances_total": STRONG_TOTAL, "instances_per_gpu": snb,
              "value": nnz * STRONG_TOTAL * args.steps / (float(tsm.item()) * 1e-3), "unit": "nnz/s",
              "ms_per_step": float(tsm.item()) / args.steps,
              "hbm_frac_per_gpu": 8 * snb * (n + m + nnz) / (float(tsm.item()) / args.steps * 1e-3) / 1e9 / hbm_peak()[0],
              "e2e": {"value": nnz * STRONG_TOTAL * e2e_steps / float(tse.item()), "unit": "nnz/s", "ms_per_step": 1e3 * float(tse.item()) / e2e_steps},
              "note": "device-timed like `value` (CUDA events, max over ranks); shards are contiguous blocks of ceil(4096/N) instances "
                      "(lpopc_b200.batch.shard_range), no data-path collective"}
    del sxs, shx

    extras = {}
    if not args.no_extras and rank == 0:
        # Lagrangian Hessian on the same batch (device-resident), reported alongside
        lam = torch.from_numpy(np.random.Generator(np.random.PCG64(2)).uniform(-1, 1, (nb, m))).to(dev)
        sg = torch.ones(nb, dtype=torch.float64, device=dev)
        d_h = torch.empty((nb, nnz_h), dtype=torch.float64, device=dev)
        for _ in range(2):
            g.eval_h_dev(nb, xs[0].data_ptr(), sg.data_ptr(), lam.data_ptr(), d_h.data_ptr())
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hs = 5
        h0.record()
        for k in range(hs):
            g.eval_h_dev(nb, xs[k % NBUF].data_ptr(), sg.data_ptr(), lam.data_ptr(), d_h.data_ptr())
        h1.record()
        torch.cuda.synchronize()
        hms = h0.elapsed_time(h1) / hs
        extras["hessian"] = {"value": nnz_h * nb / (hms * 1e-3), "unit": "nnz_h/s", "ms_per_eval": hms,
                             "bytes_per_eval": 8 * nb * (n + m + nnz_h), "kernel": "k_hess_tiled<LpbQuadrotor,6,3> + k_hess_endpoint",
                             "roofline": hbm_roofline(8 * nb * (n + m + nnz_h), hms, "hessian")}
        del lam, d_h
        # BASELINE config 5: one synthetic ns=20 / nc=6 problem on 100k LGR nodes (76 M nnz), fused eval_g + eval_jac_g
        from lpopc_b200 import examples as _ex
        g5 = nlp.TranscribedNLP(_ex.synthetic20(intervals=C5_INTERVALS, nodes=C5_NODES))
        g5.set_stream(torch.cuda.current_stream().cuda_stream)
        n5, m5, nnz5, _ = g5.get_nlp_info()
        r5 = np.random.Generator(np.random.PCG64(7))
        x5 = r5.uniform(-1.0, 1.0, n5)
        x5[-2], x5[-1] = 0.0, 10.0
        x5s = [torch.from_numpy(x5 + 1e-3 * k).to(dev) for k in range(2)]
        g5_g = torch.empty(m5, dtype=torch.float64, device=dev)
        g5_v = torch.empty(nnz5, dtype=torch.float64, device=dev)
        for k in range(3):
            g5.eval_g_jac_dev(1, x5s[k % 2].data_ptr(), g5_g.data_ptr(), g5_v.data_ptr())
        torch.cuda.synchronize()
        g5.set_option("time_kernels", 1)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c5_steps = 10
        c0.record()
        for k in range(c5_steps):
            g5.eval_g_jac_dev(1, x5s[k % 2].data_ptr(), g5_g.data_ptr(), g5_v.data_ptr())
        c1.record()
        torch.cuda.synchronize()
        c5ms = c0.elapsed_time(c1) / c5_steps
        k5ms, k5cnt = g5.kernel_time("cons_jac")
        g5.set_option("time_kernels", 0)
        # the same mesh's Lagrangian Hessian values (config 5 names both): dense dynamics, every variable pair of every row
        nnzh5 = g5.get_nlp_info()[3]
        lam5 = torch.from_numpy(r5.standard_normal(m5)).to(dev)
        sg5 = torch.ones(1, dtype=torch.float64, device=dev)
        h5_v = torch.empty(nnzh5, dtype=torch.float64, device=dev)
        g5.eval_h_dev(1, x5s[0].data_ptr(), sg5.data_ptr(), lam5.data_ptr(), h5_v.data_ptr())
        torch.cuda.synchronize()
        hc0, hc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hc0.record()
        for k in range(3):
            g5.eval_h_dev(1, x5s[k % 2].data_ptr(), sg5.data_ptr(), lam5.data_ptr(), h5_v.data_ptr())
        hc1.record()
        torch.cuda.synchronize()
        c5hms = hc0.elapsed_time(hc1) / 3
        del lam5, h5_v
        extras["c5"] = {"workload": "BASELINE config 5: synthetic ns=20/nc=6 dynamics, %d x %d = %d LGR nodes, one problem, fused eval_g+eval_jac_g"
                                    % (C5_INTERVALS, C5_NODES, C5_INTERVALS * C5_NODES),
                        "n": n5, "m": m5, "nnz_jac": nnz5, "ms_per_step": c5ms, "value": nnz5 / (c5ms * 1e-3), "unit": "nnz/s",
                        "kernel_ms": k5ms / max(1, k5cnt), "roofline": hbm_roofline(8 * (n5 + m5 + nnz5), c5ms, "c5"),
                        "hessian": {"nnz_h": nnzh5, "ms_per_eval": c5hms, "value": nnzh5 / (c5hms * 1e-3), "unit": "nnz_h/s",
                                    "kernel": "k_hess_nodes<LpbSynthetic20> (run-time pair loops: the functor set is dense, 351 variable pairs x 20 rows per node)",
                                    "roofline": hbm_roofline(8 * (n5 + m5 + nnzh5), c5hms)}}
        del g5, g5_g, g5_v, x5s

    solves = None
    if not args.no_extras:
        # batched OCP solves/s: every rank solves SOLVE_INSTANCES of its own instances to KKT tolerance with
        # the lockstep interior-point method of lpopc_b200/solver.py (all NLP callbacks device-resident)
        from lpopc_b200 import solver
        ns_ = SOLVE_INSTANCES
        ph0 = op.phases[0]
        nominal = np.array([ph0.stateguess[j][0] for j in range(len(ph0.statemin))])
        x0s = np.stack([nominal + 0.2 * np.random.Generator(np.random.PCG64(5 + first + i)).uniform(-1, 1, nominal.size) for i in range(ns_)])
        g2 = nlp.TranscribedNLP(op)
        X0 = batch.mpc_starting_points(op, g2.lgr_points(), x0s)
        g2.probe_dependencies(X0[0])  # sparse Hessian pattern, as the reference does once per problem
        ev = solver.CudaEvaluator(g2)
        if rank == 0:
            # the same Hessian evaluation on the pattern the dependency probe leaves (what the reference evaluates
            # after GetDependecies; the "hessian" entry above is the forced-dense pattern)
            nh2 = g2.get_nlp_info()[3]
            lam2 = torch.from_numpy(np.random.Generator(np.random.PCG64(2)).uniform(-1, 1, (nb, m))).to(dev)
            sg2 = torch.ones(nb, dtype=torch.float64, device=dev)
            d_h2 = torch.empty((nb, nh2), dtype=torch.float64, device=dev)
            for _ in range(2):
                g2.eval_h_dev(nb, xs[0].data_ptr(), sg2.data_ptr(), lam2.data_ptr(), d_h2.data_ptr())
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for k in range(5):
                g2.eval_h_dev(nb, xs[k % NBUF].data_ptr(), sg2.data_ptr(), lam2.data_ptr(), d_h2.data_ptr())
            p1.record()
            torch.cuda.synchronize()
            pms = p0.elapsed_time(p1) / 5
            extras["hessian_probed"] = {"value": nh2 * nb / (pms * 1e-3), "unit": "nnz_h/s", "nnz_h": nh2, "ms_per_eval": pms,
                                        "bytes_per_eval": 8 * nb * (n + m + nh2), "roofline": hbm_roofline(8 * nb * (n + m + nh2), pms)}
            del lam2, sg2, d_h2
        bxl, bxu, _, _ = ev.bounds()
        XL, XU = batch.mpc_bounds(bxl, bxu, op, x0s)
        ipm = solver.BatchedIPM(ev, tol=1e-6, max_iter=150, var_blocks=solver.interval_blocks(op, ev.n))
        ipm.solve(X0[:8], XL[:8], XU[:8])  # warm-up
        barrier()
        solve_sampler = ClockSampler(local_rank, period_s=0.015)  # back-to-back polling would steal the GIL from the solver's Python loop
        solve_sampler.start()
        ts = time.perf_counter()
        _prof = None
        if os.environ.get("LPB_BENCH_PROFILE_SOLVE"):  # development: where does the solve spend its host time?
            import cProfile
            _prof = cProfile.Profile()
            _prof.enable()
        res = ipm.solve(X0, XL, XU, chunk=SOLVE_INSTANCES)
        torch.cuda.synchronize()
        if _prof is not None:
            import pstats
            _prof.disable()
            pstats.Stats(_prof, stream=sys.stderr).sort_stats("tottime").print_stats(16)
        t_own = time.perf_counter() - ts  # this rank's own solve time (the barrier below waits for the slowest rank)
        barrier()
        t_solve = torch.tensor([time.perf_counter() - ts], dtype=torch.float64, device=dev)
        per_rank = torch.zeros((world, 2), dtype=torch.float64, device=dev)  # [seconds, lockstep iterations] of every rank
        per_rank[rank, 0], per_rank[rank, 1] = t_own, float(res["iters"].max().item())
        solve_clocks = solve_sampler.stop()
        n_ok = (res["status"] == 0).sum().to(torch.float64).reshape(1)
        it_sum = res["iters"].sum().to(torch.float64).reshape(1)
        if world > 1:
            dist.all_reduce(t_solve, op=dist.ReduceOp.MAX)
            dist.all_reduce(n_ok, op=dist.ReduceOp.SUM)
            dist.all_reduce(it_sum, op=dist.ReduceOp.SUM)
            dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        solves = {"metric": "batched OCP solves/s", "value": float(n_ok.item()) / float(t_solve.item()), "unit": "solves/s",
                  "instances": ns_ * world, "converged": int(n_ok.item()), "seconds": float(t_solve.item()),
                  "iters_mean": float(it_sum.item()) / (ns_ * world), "tol": 1e-6, "nnz_h_probed": ev.nnz_h,
                  "clocks_rank0": {k: solve_clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")},
                  "seconds_per_rank": [round(v, 3) for v in per_rank[:, 0].tolist()], "iters_max_per_rank": [int(v) for v in per_rank[:, 1].tolist()],
                  "solver": "lockstep primal-dual interior point, exact FD Hessian, KKT step: %s; block solves = lpb_blocktri_solve "
                            "(one launch per solve), factorisation = %s; NLP callbacks = device-resident transcription kernels"
                            % (ipm.kkt_kind, "lpb_kkt_factor (assembly incl. gamma J^T J, block Cholesky and inertia retry in one launch per iteration)"
                               if getattr(ipm.kkt, "fused_factor", False) and not getattr(ipm.kkt, "_kkt_fused_off", False)
                               else "cuSOLVER batched Cholesky + cuBLAS")}
        del g2, ev, ipm, res

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        ph = op.phases[0]
        N, ns, nc, npth = ph.GetTotalNodes(), len(ph.statemin), len(ph.controlmin), len(ph.pathmin)
        # k_cons_jac per instance: reads x (n), writes the node rows of g, the node part of NL and the
        # constant segment C (ns copies of the Doffdiag values, fused into the kernel)
        kbytes = 8 * nb * (n + (ns + npth) * N + (ns + npth) * (ns + nc + 2) * N + ns * sum(int(v) ** 2 for v in ph.nodesperinterval))
        kavg_ms = kern_ms / max(1, kern_cnt)
        achieved = kbytes / (kavg_ms * 1e-3) / 1e9 if kavg_ms > 0 else 0.0
        # CPU baseline on a bounded sample, same box, all host threads
        threads = os.cpu_count() or 1
        # mesh-constant tail [L | C] of the values: 2 per linear row + ns * sum_k N_k^2 Doffdiag entries
        const_tail = 2 * (len(op.phases) + len(op.links)) + ns * sum(int(v) ** 2 for v in ph.nodesperinterval)
        line = {
            "metric": "collocation Jacobian nnz/s", "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(_Sizes(n, m, nnz, nnz_h), world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "nnz/s", "h2d_bytes_per_step": 8 * nb * n,
                    "d2h_bytes_per_step": 8 * nb * (m + (sparse_on if sparse_calls else nnz - const_tail)),
                    "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps, "host_buffers": "pinned", "cpu_affinity": affinity,
                    "sparse_return": {"calls": sparse_calls, "values_sent_per_instance": sparse_on, "head_values_per_instance": nnz - const_tail,
                                      "segments_refetched": sparse_fixups},
                    "persistent_values": {"enabled": not args.no_persistent, "calls_without_host_writes": persistent_hits},
                    "pcie_floor": {"d2h_bytes": d2h_bytes, "ms": pcie_ms, "gbs": d2h_bytes / (pcie_ms * 1e-3) / 1e9, "ms_rank0": pcie_ms_rank,
                                   "e2e_over_floor": 1e3 * float(te.item()) / e2e_steps / pcie_ms,
                                   "how": "one pinned D2H cudaMemcpy per rank of the bytes the call returns, with the H2D of x on a second "
                                          "stream; ALL ranks copy at the same time (barrier first), CUDA events, max over ranks, after the "
                                          "timed e2e region: the host-side ceiling the e2e number scales against"},
                    "note": "all nnz_jac values are delivered per call and checked against the device-resident evaluation; the %d "
                            "mesh-constant values per instance (linear rows + Doffdiag segment) are written into the caller's array by host "
                            "threads from a cached copy, and of the %d x-dependent values only the (row block, column block) segments that "
                            "are not all-zero cross PCIe (the rest is verified to be zero on the device every call and zero-filled by the "
                            "host threads)" % (const_tail, nnz - const_tail)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_cons_jac<LpbQuadrotor,WANT_G=1,WANT_JAC=1,UNROLL=1>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES, "traffic_source": NCU_TRAFFIC_SOURCE, "bytes_per_launch": kbytes, "avg_launch_ms": kavg_ms,
                         "launches_timed": kern_cnt, "peak_source": peak_src,
                         "step_bytes": 8 * nb * (n + m + nnz), "step_frac": 8 * nb * (n + m + nnz) / (ms / args.steps * 1e-3) / 1e9 / peak},
            "cpu_baseline": (cpu[3].baseline(cpu[0], threads, "all %d quadrotor instances x %d passes in %.1f s; %s" % (nb, cpu[2], cpu[1], cpu[3].describe()))
                             if cpu else {"value": None, "unit": "nnz/s", "cores": threads, "kind": "skipped (--no-cpu)", "sample": ""}),
            "objective_checksum": float(np.sum(all_f)),
        }
        line.update(extras)
        line["strong"] = strong
        if solves:
            line["solves"] = solves
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
