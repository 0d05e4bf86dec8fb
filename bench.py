#!/usr/bin/env python
"""bench.py — batch node-evaluation throughput (Mnodes/s) on B200, BASELINE.json's metric.

A *step* is one pass of the hot path over one batch of synthetic parent nodes:
  * headline workload (BASELINE.json configs[1]): N-Queens N=17, g=1 — a batch of --M parents whose depths
    follow the explored-tree depth histogram of the N=17 search (what the reference's step-2 loop offloads),
    boards = random conflict-free prefixes + random remaining order;
  * secondary workload reported in the same line under "pfsp": PFSP ta014, lb1, ub=1 — parents with the
    ta014/lb1 offload depth histogram, random permutations.
`value`  = parents evaluated per second, inputs and outputs resident in HBM (device entry point, one kernel
           launch per step, CUDA events on the launching stream, max over ranks);
`e2e`    = the same metric through the reference-facing C-ABI call tsb_nq_evaluate / tsb_pfsp_evaluate with
           HOST buffers: the copy of the chunk to the device and of the labels/bounds back happen inside the
           timed region, every step;
`roofline` = algorithmic bytes per launch (21 B + N B per parent; 88 B + 80 B for PFSP) / measured kernel time,
           against MEASURED_PEAKS.json's HBM copy bandwidth;
`at_M50000` = the same three numbers at the reference's default chunk size --M 50000 (launch-latency bound:
           1.9 MB per launch), which is what one offload of the unmodified Chapel driver would see.
`search`   = whole searches (explored tree / wall time, the quantity the reference prints) with the pool of
           step 2 resident in HBM (tsb_*_search_device: count + build kernels per round, nothing but three
           counters crosses PCIe): N-Queens N=17 at --M and at the reference's --M 50000, PFSP ta014/lb1 and
           ta020/lb2 at --M 50000 (BASELINE configs 2-4).  Counts are checked against the reference's.  With
           N > 1 ranks every rank runs one task of the reference's static N-way split of the warm-up pool
           (tsb_*_search_device_part) on its own GPU; tree = sum over ranks, time = max over ranks ("strong").
Between timed iterations the inputs/outputs rotate over buffer sets whose total footprint exceeds the 126 MB L2.

`--impl reference` times the reference's own CPU implementation of the same path (its isSafe / lb1_bound
loops, compiled unmodified into oracle/_ref; the oracle port if that is absent) on all host threads.

N > 1 (torchrun): every rank evaluates its own batch on its own GPU (the path shards by parents, no
collective in the data path); value = total parents / max-over-ranks time; scaling = weak.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gpu-accelerated-tree-search-chapel_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

L2_BYTES = 126 * 2**20

# explored-tree nodes per depth of the N=17 search = depth histogram of the parents its offload loop evaluates
# (exact counts, tests/golden/nqueens_depth_hist.json; sum over depths 1..17 = 8 017 021 931 = exploredTree)
_HIST_PATH = os.path.join(ROOT, "tests", "golden", "nqueens_depth_hist.json")
# depth histogram of offloaded parents, ta014 lb1 (depth: count)
PFSP_TA014_LB1_HIST = {1: 1, 2: 39, 3: 165, 4: 639, 5: 2252, 6: 7003, 7: 19626, 8: 50445, 9: 116575, 10: 231523,
                       11: 366493, 12: 464303, 13: 466099, 14: 366546, 15: 240078, 16: 144463, 17: 72917,
                       18: 21829, 19: 2648}


# depth histogram of offloaded parents, ta020 lb2
PFSP_TA020_LB2_HIST = {1: 8, 2: 92, 3: 597, 4: 3538, 5: 17546, 6: 71163, 7: 224227, 8: 533085, 9: 947932, 10: 1193333,
                       11: 1044659, 12: 572606, 13: 205065, 14: 47233, 15: 7546, 16: 1526, 17: 208, 18: 13}


# ----------------------------------------------------------------------------------------- synthetic inputs
def nq_depth_hist(N):
    with open(_HIST_PATH) as f:
        h = json.load(f)[str(N)]
    return {int(k): int(v) for k, v in h.items()}


def synth_nq_parents(N, count, seed, dtype):
    """parents as the reference's pool holds them: board[0..depth) a conflict-free placement (random walk of
    the search tree), board[depth..N) the remaining values in random order; depth ~ explored-tree histogram"""
    rng = np.random.default_rng(seed)
    hist = nq_depth_hist(N)
    depths = np.array(sorted(hist), dtype=np.int64)
    prob = np.array([hist[int(d)] for d in depths], dtype=np.float64)
    base = min(count, 1 << 18)  # distinct random walks; replicated by random gather up to `count`
    target = rng.choice(depths, size=base, p=prob / prob.sum())
    board = np.zeros((base, 20), dtype=np.uint8)
    todo = np.arange(base)
    full = (1 << N) - 1
    while todo.size:
        n = todo.size
        cols = np.zeros(n, dtype=np.int64)
        ld = np.zeros(n, dtype=np.int64)
        rd = np.zeros(n, dtype=np.int64)
        ok = np.ones(n, dtype=bool)
        tgt = target[todo]
        rows = np.zeros((n, N), dtype=np.uint8)
        for r in range(int(tgt.max())):
            act = ok & (tgt > r)
            free = ~(cols | ld | rd) & full
            act &= free != 0
            ok &= (tgt <= r) | (free != 0)
            # choose a uniformly random set bit of `free`
            cnt = np.zeros(n, dtype=np.int64)
            for b in range(N):
                cnt += (free >> b) & 1
            pick = (rng.random(n) * np.maximum(cnt, 1)).astype(np.int64)
            chosen = np.zeros(n, dtype=np.int64)
            seen = np.zeros(n, dtype=np.int64)
            for b in range(N):
                bit = (free >> b) & 1
                hit = (bit == 1) & (seen == pick)
                chosen = np.where(hit, b, chosen)
                seen += bit
            bitv = np.where(act, 1 << chosen, 0)
            rows[:, r] = np.where(act, chosen, 0)
            cols |= bitv
            ld = ((ld | bitv) << 1) & full
            rd = (rd | bitv) >> 1
        done = ok
        idx = todo[done]
        # remaining values in random order after the placed prefix
        for i, src in zip(idx, np.nonzero(done)[0]):
            d = int(target[i])
            placed = rows[src, :d]
            rest = np.setdiff1d(np.arange(N, dtype=np.uint8), placed)
            board[i, :d] = placed
            board[i, d:N] = rng.permutation(rest)
        todo = todo[~done]
    sel = rng.integers(0, base, size=count) if count > base else np.arange(count)
    out = np.zeros(count, dtype=dtype)
    out["depth"] = target[sel]
    out["board"] = board[sel]
    return out


def synth_pfsp_parents(count, seed, dtype, hist=PFSP_TA014_LB1_HIST, jobs=20):
    rng = np.random.default_rng(seed)
    depths = np.array(sorted(hist), dtype=np.int64)
    prob = np.array([hist[int(d)] for d in depths], dtype=np.float64)
    d = rng.choice(depths, size=count, p=prob / prob.sum())
    out = np.zeros(count, dtype=dtype)
    out["depth"] = d
    out["limit1"] = d - 1
    out["prmu"][:, :jobs] = np.argsort(rng.random((count, jobs)), axis=1).astype(np.int32)
    return out


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples SM clock and throttle reasons of one GPU while a timed region runs (NVML)"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "sw_power_cap": 0x4, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------- distributed helpers
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def dist_init(backend):
    import torch.distributed as dist
    _, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            import torch
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, device_id=torch.device(f"cuda:{local_rank}"))
        else:
            dist.init_process_group(backend=backend)
    return world


def dist_barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def dist_max_sum(world, t_seconds, units, device):
    """(max over ranks of t_seconds, sum over ranks of units)"""
    if world <= 1:
        return t_seconds, units
    import torch
    import torch.distributed as dist
    t = torch.tensor([t_seconds], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


# ----------------------------------------------------------------------------------------- our arm
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_workload(kind, M, steps, warmup, device_index, world, N=17):
    """returns dict with value / e2e / kernel timing for one workload at chunk size M"""
    import torch

    import tsb200
    dev = torch.device(f"cuda:{device_index}")
    torch.cuda.set_device(dev)
    rank, _, _ = dist_env()
    if kind == "nq":
        in_rec, out_rec = 21, N
        make = lambda seed: synth_nq_parents(N, M, seed, tsb200.NQ_NODE_DTYPE)  # noqa: E731
        ev = tsb200.NQueensEvaluator(N, 1, M, device=device_index)
        call_dev = lambda i, o, s: ev.evaluate_device(i, M, o, s)  # noqa: E731
        out_dtype, out_elems = np.uint8, M * N
        call_host = lambda par, out: ev.evaluate_gpu(par, M * N, out)  # noqa: E731
    else:
        in_rec, out_rec = 88, 80
        inst, lb, best, hist = (14, "lb1", 1377, PFSP_TA014_LB1_HIST) if kind == "pfsp" else \
            (20, "lb2", 1591, PFSP_TA020_LB2_HIST)
        make = lambda seed: synth_pfsp_parents(M, seed, tsb200.PFSP_NODE_DTYPE, hist)  # noqa: E731
        ev = tsb200.PfspEvaluator(inst, M=M, device=device_index)
        call_dev = lambda i, o, s: ev.evaluate_device(lb, i, M, best, o, s)  # noqa: E731
        out_dtype, out_elems = np.int32, M * 20
        call_host = lambda par, out: ev.evaluate_gpu(par, M * 20, best, lb, out)  # noqa: E731
    bytes_per_set = M * (in_rec + out_rec)
    nsets = max(2, int(np.ceil(2.5 * L2_BYTES / bytes_per_set)))  # rotate over > 2.5x L2 of distinct buffers
    nsets = min(nsets, 64)
    base = make(1234 + rank)
    host_in = [base] + [np.roll(base, 7919 * (k + 1), axis=0) for k in range(min(nsets, 4) - 1)]
    d_in = [torch.from_numpy(host_in[k % len(host_in)].view(np.uint8).reshape(-1)).to(dev) for k in range(nsets)]
    d_out = [torch.empty(M * out_rec, dtype=torch.uint8, device=dev) for _ in range(nsets)]
    # a real (non-default) stream: the kernels are launched on it and the CUDA events are recorded on it
    tstream = torch.cuda.Stream(device=dev)
    stream = tstream.cuda_stream
    assert stream != 0

    # ---- device-resident: one launch per step
    torch.cuda.synchronize()
    with torch.cuda.stream(tstream):
        for w in range(warmup):
            call_dev(d_in[w % nsets].data_ptr(), d_out[w % nsets].data_ptr(), stream)
    torch.cuda.synchronize()
    l0 = ev.kernel_launches
    dist_barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device_index) as clk, torch.cuda.stream(tstream):
        e0.record(tstream)
        for k in range(steps):
            call_dev(d_in[k % nsets].data_ptr(), d_out[k % nsets].data_ptr(), stream)
        e1.record(tstream)
        torch.cuda.synchronize()
    dist_barrier(world)
    t_dev = e0.elapsed_time(e1) / 1e3
    launches = ev.kernel_launches - l0
    t_dev_max, units = dist_max_sum(world, t_dev, M * steps, dev)

    # ---- end to end through the host-buffer C-ABI call (H2D + kernel + D2H inside the timed region)
    host_out = [np.empty(out_elems, dtype=out_dtype) for _ in range(len(host_in))]
    for w in range(max(warmup, len(host_in))):  # first calls page-lock the caller's arrays
        call_host(host_in[w % len(host_in)], host_out[w % len(host_in)])
    dist_barrier(world)
    torch.cuda.synchronize()
    with ClockSampler(device_index) as clk2:  # (the device-resident region above lasts only a few ms)
        t0 = time.perf_counter()
        for k in range(steps):
            call_host(host_in[k % len(host_in)], host_out[k % len(host_in)])
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
    clk.samples += clk2.samples
    clk.reasons |= clk2.reasons
    dist_barrier(world)
    t_e2e_max, _ = dist_max_sum(world, t_e2e, M * steps, dev)
    ev.close()
    del d_in, d_out
    torch.cuda.empty_cache()
    return {
        "M": M, "steps": steps, "units": units, "t_dev": t_dev_max, "t_dev_local": t_dev, "t_e2e": t_e2e_max,
        "launches": launches, "in_rec": in_rec, "out_rec": out_rec, "nsets": nsets, "clocks": clk.summary(),
        "footprint_mb": nsets * bytes_per_set / 2**20, "sample": base,
    }


# ----------------------------------------------------------------------------------------- whole searches
GOLDEN_TREES = {("nq", 17): (8017021931, 95815104), ("pfsp", 14, "lb1"): (2573652, 2648, 1377),
                ("pfsp", 20, "lb2"): (4870386, 0, 1591)}  # tests/golden/counts.json (reference binaries)


def run_search(kind, world, rank, device_index, reps, **kw):
    """best-of-`reps` whole search; rank r runs task r of the world-way static split on its own GPU"""
    import torch

    import tsb200
    dev = torch.device(f"cuda:{device_index}")
    best_t, out = None, None
    for _ in range(reps):
        dist_barrier(world)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if kind == "nq":
            st = tsb200.nqueens_search_device_part(kw["N"], 1, 25, kw["M"], world, rank, device_index)
        else:
            st = tsb200.pfsp_search_device_part(kw["inst"], kw["lb"], 1, 25, kw["M"], world, rank, device_index)
        dt = time.perf_counter() - t0
        t_max, tree = dist_max_sum(world, dt, st.explored_tree, dev)
        _, sol = dist_max_sum(world, 0.0, st.explored_sol, dev)
        _, offl = dist_max_sum(world, 0.0, st.offloads, dev)
        _, launches = dist_max_sum(world, 0.0, st.kernel_launches, dev)
        if best_t is None or t_max < best_t:
            best_t = t_max
            out = {"explored_tree": int(tree), "explored_sol": int(sol), "seconds": t_max,
                   "value": tree / t_max / 1e6, "unit": "Mnodes/s", "offloads": int(offl),
                   "gpu_launches": int(launches), "M": kw["M"], "reps": reps, "rank0_share": st.explored_tree / tree}
            if kind == "pfsp":
                out["optimum"] = int(st.best)
    key = ("nq", kw["N"]) if kind == "nq" else ("pfsp", kw["inst"], kw["lb"])
    if key in GOLDEN_TREES:
        want = GOLDEN_TREES[key]
        out["counts_match_reference"] = (out["explored_tree"], out["explored_sol"]) == want[:2]
    return out


def summarize(r, peak, peak_src, traffic=None):
    alg_bytes = r["M"] * (r["in_rec"] + r["out_rec"])
    t_kernel = r["t_dev_local"] / r["steps"]
    achieved = alg_bytes / t_kernel / 1e9
    return {
        "value": r["units"] / r["t_dev"] / 1e6,
        "ms_per_step": r["t_dev"] / r["steps"] * 1e3,
        "e2e": {"value": r["units"] / r["t_e2e"] / 1e6, "unit": "Mnodes/s",
                "h2d_bytes_per_step": r["M"] * r["in_rec"], "d2h_bytes_per_step": r["M"] * r["out_rec"],
                "ms_per_step": r["t_e2e"] / r["steps"] * 1e3},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": alg_bytes,
                     "kernel_us": t_kernel * 1e6},
    }


# ----------------------------------------------------------------------------------------- CPU arms
def cpu_eval(kind, parents, threads, N=17, repeat=1):
    """times the reference's CPU implementation of the path on `threads` host threads (ctypes drops the GIL);
    returns (seconds, kind_string)"""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor

    from oracle import pyoracle as po
    P = parents.shape[0]
    cuts = np.linspace(0, P, threads + 1).astype(int)
    if kind == "nq":
        out = np.zeros(P * N, dtype=np.uint8)
        if po.ref_available():
            f = po.ref_nqueens().ref_nq_evaluate_range_rep
            work = lambda a, b: f(parents.ctypes.data, int(a), int(b), N, 1, out.ctypes.data, repeat)  # noqa: E731
            src = "reference"
        else:
            f = po.lib().or_nq_evaluate_range
            work = lambda a, b: [f(parents.ctypes.data, int(a), int(b), N, 1, out.ctypes.data)  # noqa: E731
                                 for _ in range(repeat)]
            src = "port"
    else:
        out = np.zeros(P * 20, dtype=np.int32)
        if po.ref_available():
            d1, d2 = po.ref_pfsp_data(14)
            f = po.ref_pfsp().ref_pfsp_evaluate_range_rep
            work = lambda a, b: f(d1, d2, 1, parents.ctypes.data, int(a), int(b), 1377, out.ctypes.data, repeat)  # noqa: E731
            src = "reference"
        else:
            t = po.tables(14)
            f = po.lib().or_pfsp_evaluate_range
            work = lambda a, b: [f(C.byref(t), 1, parents.ctypes.data, int(a), int(b), 1377, out.ctypes.data)  # noqa: E731
                                 for _ in range(repeat)]
            src = "port"
    def job(ab):  # every thread sweeps its slice `repeat` times inside ONE foreign call (GIL released)
        work(*ab)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(job, zip(cuts[:-1], cuts[1:])))  # warm: threads started, pages touched
        t0 = time.perf_counter()
        list(ex.map(job, zip(cuts[:-1], cuts[1:])))
        dt = time.perf_counter() - t0
    return dt / repeat, src


def cpu_baseline(kind, sample, threads):
    """~10-20 s of CPU work in total: the slice of each thread is swept `repeat` times"""
    sample = np.ascontiguousarray(sample)
    per_thread_rate = 12e6 if kind == "nq" else 0.7e6  # parents/s/thread, order of magnitude
    repeat = max(1, int(15.0 * per_thread_rate * threads / sample.shape[0] / threads))
    dt, src = cpu_eval(kind, sample, threads, repeat=repeat)
    return {"value": sample.shape[0] / dt / 1e6, "unit": "Mnodes/s", "cores": threads, "kind": src,
            "sample": f"{sample.shape[0]} parents of the step's batch x {repeat} sweeps, {dt * repeat:.2f} s wall"}


def emit(line):
    """print THE one JSON line on the real stdout (fd 1 is pointed at stderr while the bench runs, so that
    libraries that write to stdout on their own — NCCL prints its version there — cannot pollute it)"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--M", type=int, default=1 << 22, help="parents per step and per GPU (the drivers' --M)")
    ap.add_argument("--pfsp-M", type=int, default=1 << 20)
    ap.add_argument("--no-pfsp", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the --M 50000 measurements")
    ap.add_argument("--lb2", action="store_true", help="also time PFSP ta020 lb2 (BASELINE configs[3])")
    ap.add_argument("--no-search", action="store_true", help="skip the whole-search measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, local_rank, world = dist_env()
    N = 17
    config = {"workload": f"N-Queens N={N} g=1 batch node evaluation (BASELINE configs[1]); synthetic parents, "
                          f"depth ~ explored-tree histogram of the N={N} search, conflict-free prefixes",
              "N": N, "g": 1, "M": args.M, "M_note": "parents per step per GPU = the drivers' --M; the reference "
              "default --M 50000 is reported under at_M50000", "parallelism": f"{args.gpus} x independent pools",
              "l2": "inputs/outputs rotate over buffer sets totalling > 2.5x the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        import tsb200  # dtype only
        threads = os.cpu_count() or 1
        P = min(args.M, 1 << 21)
        sample = synth_nq_parents(N, P, 1234, tsb200.NQ_NODE_DTYPE)
        # bounded: the whole run (warm-up + steps) is sized for about one minute of wall time on this host
        per_step_s = min(2.0, 60.0 / (args.steps + args.warmup))
        rep = max(1, int(per_step_s * 1.3e6 * threads / P))  # ~1.3 M parents/s/thread for the reference's isSafe loop
        for _ in range(args.warmup):
            cpu_eval("nq", sample, threads, repeat=1)
        t, src = 0.0, "port"
        for _ in range(args.steps):
            dt, src = cpu_eval("nq", sample, threads, repeat=rep)
            t += dt
        v = P * args.steps / t / 1e6
        line = {"impl": "reference", "metric": "Mnodes/s", "value": v, "unit": "Mnodes/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "Mnodes/s", "cores": threads, "kind": src,
                                 "sample": f"{P} parents per step (bounded sample of the --M {args.M} batch)"},
                "e2e": {"value": v, "unit": "Mnodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if not args.no_pfsp:
            Pp = 1 << 17
            ps = synth_pfsp_parents(Pp, 99, tsb200.PFSP_NODE_DTYPE)
            dt, src2 = cpu_eval("pfsp", ps, threads, repeat=max(1, int(3.0 * 0.08e6 * threads / Pp)))
            line["pfsp"] = {"workload": "PFSP ta014 lb1 ub=1", "value": Pp / dt / 1e6, "unit": "Mnodes/s",
                            "cores": threads, "kind": src2, "sample": f"{Pp} parents"}
        emit(line)
        return

    import torch
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    world = dist_init("nccl")
    device_index = local_rank if world > 1 else 0
    peak, peak_src = peaks()

    big = run_workload("nq", args.M, args.steps, args.warmup, device_index, world, N)
    # DRAM bytes of one launch from the ncu --set full capture of this very launch shape (N=17, 4 Mi parents:
    # profiles/nq_eval_r1_ncu.txt, dram__bytes_read.sum + dram__bytes_write.sum; the labels written last are
    # still in the 126 MB L2 when the kernel ends)
    traffic = 88130560 + 14575360 if (args.M == 1 << 22 and N == 17) else None
    main_s = summarize(big, peak, peak_src, traffic=traffic)
    line = {"metric": "Mnodes/s", "value": main_s["value"], "unit": "Mnodes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_s["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "clocks": big["clocks"], "e2e": main_s["e2e"], "gpu_launches": big["launches"],
            "roofline": main_s["roofline"]}
    line["config"]["buffer_sets"] = big["nsets"]
    line["config"]["footprint_mb"] = round(big["footprint_mb"], 1)
    if not args.no_small:
        small = run_workload("nq", 50000, max(args.steps * 10, 200), args.warmup, device_index, world, N)
        s = summarize(small, peak, peak_src)
        line["at_M50000"] = {"value": s["value"], "ms_per_step": s["ms_per_step"], "e2e": s["e2e"],
                             "roofline": s["roofline"], "note": "one offload of the unmodified driver: 1.9 MB per "
                             "launch, launch-latency bound by construction (SURVEY.md hard part 2)"}
    if not args.no_pfsp:
        pf = run_workload("pfsp", args.pfsp_M, args.steps, args.warmup, device_index, world)
        # DRAM bytes of one launch of this shape (ta014, 1 Mi parents): profiles/pfsp_lb1_r1_ncu.txt
        ps = summarize(pf, peak, peak_src, traffic=92355072 + 27396096 if args.pfsp_M == 1 << 20 else None)
        line["pfsp"] = {"workload": "PFSP ta014 lb1 ub=1, synthetic parents with the ta014/lb1 offload depth histogram",
                        "M": args.pfsp_M, "value": ps["value"], "unit": "Mnodes/s", "ms_per_step": ps["ms_per_step"],
                        "e2e": ps["e2e"], "roofline": ps["roofline"], "gpu_launches": pf["launches"],
                        "dtype": "int32"}
        if not args.no_small:
            pf2 = run_workload("pfsp", 50000, max(args.steps * 5, 100), args.warmup, device_index, world)
            s2 = summarize(pf2, peak, peak_src)
            line["pfsp"]["at_M50000"] = {"value": s2["value"], "e2e": s2["e2e"], "roofline": s2["roofline"]}
    if args.lb2:
        l2 = run_workload("lb2", 1 << 18, max(3, args.steps // 10), args.warmup, device_index, world)
        ls = summarize(l2, peak, peak_src, traffic=23106048 + 512)  # profiles/pfsp_lb2_r1_ncu.txt (bounds stay in L2)
        line["pfsp_lb2"] = {"workload": "PFSP ta020 lb2 ub=1 (best=1591 at launch), synthetic parents with the "
                            "ta020/lb2 offload depth histogram; int-ALU bound (O(pairs*jobs) per child)",
                            "M": 1 << 18, "value": ls["value"], "unit": "Mnodes/s", "ms_per_step": ls["ms_per_step"],
                            "e2e": ls["e2e"], "roofline": ls["roofline"], "gpu_launches": l2["launches"]}
    if not args.no_search:
        reps = 3
        srch = {"note": "explored tree / wall time of whole searches, pool resident in HBM (tsb_*_search_device); "
                        "time includes step 1 and 3 on the CPU, handle creation and arena allocation",
                "scaling": "strong" if world > 1 else None,
                "nqueens_N17": run_search("nq", world, rank, device_index, reps, N=17, M=args.M),
                "nqueens_N17_M50000": run_search("nq", world, rank, device_index, 1, N=17, M=50000),
                # (millisecond-scale searches: more repetitions, a sporadic slow cudaMalloc is tens of ms)
                "pfsp_ta014_lb1_M50000": run_search("pfsp", world, rank, device_index, 7, inst=14, lb="lb1", M=50000),
                "pfsp_ta020_lb2_M50000": run_search("pfsp", world, rank, device_index, 7, inst=20, lb="lb2", M=50000)}
        line["search"] = srch
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        line["cpu_baseline"] = cpu_baseline("nq", big["sample"][: 1 << 21], threads)
        one = cpu_baseline("nq", big["sample"][: 1 << 19], 1)
        line["cpu_baseline"]["value_1core"] = one["value"]
        if not args.no_pfsp:
            line["pfsp"]["cpu_baseline"] = cpu_baseline("pfsp", pf["sample"][: 1 << 17], threads)
    if rank == 0:
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
