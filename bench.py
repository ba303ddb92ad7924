#!/usr/bin/env python
"""bench.py -- SG operator apply throughput (GDoF/s = N*|Lambda| per apply) on 1..8 B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A step is one application y = A w of the stochastic-Galerkin operator to one slab.

  value      whole-job GDoF/s with the input slab already resident in HBM (CUDA events, max over
             ranks, barrier + synchronize on both sides)
  e2e        same metric through the public API with HOST buffers: pinned host slab -> H2D ->
             MultiOperator.apply -> D2H of the result, every step (copies of neighbouring steps overlap
             the kernel on separate streams; the PCIe link bounds it)
  roofline   algorithmic (compulsory) bytes of one apply / kernel time vs the measured HBM peak
  cpu_baseline  the oracle port of the reference's CPU path (dict of vectors, one scipy SpMV per
             (mu, m)) timed on a bounded sample of output columns of the same workload
  --impl reference   times that CPU path as the main line (the reference itself is Python 2 and
             cannot be imported here)

Workloads (BASELINE.json configs, synthetic inputs of SURVEY.md 8d):
  c3  (default) 1000 x 1000 grid (N = 10^6), M = 50, anisotropic Lambda with 988 indices, Hermite
      chaos -- the configuration the headline target (N = 1M x |Lambda| = 1000) is quoted on
  c2  512 x 512 grid, M = 20, total degree 3 (1771 indices), Legendre
  c4  500 x 500 grid, M = 30, first 5000 indices of the degree-3 set, Legendre
  c1  32 x 32 grid, M = 5, total degree 2 (21 indices): the reference's CPU-runnable case
The workload is the same for every N (total work fixed) => "scaling": "strong".  Default `--shard rows`: the grid rows are
sharded over the ranks (every rank owns N/G spatial dofs of all |Lambda| columns); one step = exchange
of the two boundary grid rows with the neighbouring ranks (point to point over NCCL) + the single-GPU
kernel on the local grid.  `--shard m` is the decomposition BASELINE.json's north_star describes: the M
expansion terms are sharded (balanced by term count), every rank applies its terms to a full replica
of w and the partial results are combined with an NCCL reduce-scatter along Lambda; it moves a whole
slab over NVLink per step and cannot scale against a fused kernel.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    "c1": dict(n=32, M=5, lam=("td", 5, 2), kind="legendre", mu=0.0),
    "c2": dict(n=512, M=20, lam=("td", 20, 3), kind="legendre", mu=0.0),
    "c3": dict(n=1000, M=50, lam=("aniso", 50, 11.3), kind="hermite", mu=0.0),
    "c4": dict(n=500, M=30, lam=("prefix", 30, 3, 5000), kind="legendre", mu=0.0),
    # config 3 on an "unstructured" numbering: the same blocks with the dofs renumbered tile by tile (16 x 16
    # tiles, random order inside a tile), so that the stencil recognition rejects the pattern and the
    # general sparse path (k_apply_ellblk) runs
    "c3u": dict(n=1000, M=50, lam=("aniso", 50, 11.3), kind="hermite", mu=0.0, perm=16),
    "c2u": dict(n=512, M=20, lam=("td", 20, 3), kind="legendre", mu=0.0, perm=16),
    # config 3 in its natural (row-major) numbering but pushed through the general sparse path: what a mesh
    # with a bandwidth-reducing numbering looks like to that kernel
    "c3n": dict(n=1000, M=50, lam=("aniso", 50, 11.3), kind="hermite", mu=0.0, path="ell"),
    # c3 on a 256 x 256 grid: same Lambda and schedule, 15x less work -- for ncu captures only
    "c5t": dict(n=64, M=8, lam=("aniso", 8, 9.0), kind="hermite", mu=0.0),
    "c3t": dict(n=64, M=50, lam=("aniso", 50, 11.3), kind="hermite", mu=0.0),
    "c3s": dict(n=256, M=50, lam=("aniso", 50, 11.3), kind="hermite", mu=0.0),
}


def workload_name(key):
    c = CONFIGS[key]
    lam = c["lam"]
    if lam[0] == "td":
        ls = "total-degree-%d Lambda" % lam[2]
    elif lam[0] == "aniso":
        ls = "anisotropic downward-closed Lambda (T=%s)" % lam[2]
    else:
        ls = "first %d indices of the degree-%d set" % (lam[3], lam[2])
    extra = ", dofs renumbered in random order inside %dx%d tiles (general sparse path)" % (c["perm"], c["perm"]) if c.get("perm") else ""
    if c.get("path") == "ell":
        extra = ", natural numbering forced through the general sparse path"
    return "%s: 2D diffusion %dx%d FD grid (N=%d), M=%d, %s, %s chaos%s" % (
        key, c["n"], c["n"], c["n"] ** 2, c["M"], ls, c["kind"], extra)


def bench_config(key, N, L, M):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm."""
    return {"workload": workload_name(key), "N": N, "L": L, "M": M,
            "l2_note": "input slab %.1f GB >> 126 MB L2, no flush needed" % (N * L * 8 / 1e9)}


def index_set(cfg):
    from spuq_b200 import synthetic
    lam = cfg["lam"]
    if lam[0] == "td":
        return synthetic.complete_order_indices(lam[1], lam[2])
    if lam[0] == "aniso":
        return synthetic.anisotropic_indices(lam[1], lam[2])
    return synthetic.truncated_complete_indices(lam[1], lam[2], lam[3])


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- CPU path (oracle port of the reference) ------------------------------------------------------
class CpuReference:
    """The reference-faithful CPU path on a sample of output columns.

    The reference's apply is a loop over mu (multi_operator2.py:52-83); output column mu costs one
    leaf apply of A_0 plus, for every m < maxm, the vector combination and one leaf apply of A_m,
    independently of all other columns, so the time of a sample of columns extrapolates linearly
    to |Lambda|.  The timed loop is oracle.sg.MultiOperator.apply restricted to the sampled
    columns (``only=``): dict-of-vectors MultiVector, Multiindex `in Lambda` scans, one scipy SpMV
    per (mu, m), cached CSR leaves -- single-threaded, as the reference is.

    Everything on this side (index set, neighbour look-ups, beta values, blocks) is built by the
    oracle package alone (oracle/synth.py, oracle/sampled.py): the product library is not loaded."""

    def __init__(self, cfg, n_cols, seed=1234, get_column=None):
        import numpy as np
        from oracle import synth
        from oracle.sampled import SampledApply
        n, M = cfg["n"], cfg["M"]
        self.N = N = n * n
        idx = synth.index_set(cfg["lam"])
        self.L = L = idx.shape[0]
        mats = synth.fd_matrices(n, M)
        if cfg.get("perm"):
            mats = synth.permuted(mats, synth.tile_permutation(n, cfg["perm"]))
        pick = np.unique(np.linspace(0, L - 1, min(n_cols, L)).astype(int))
        rng = np.random.RandomState(seed)
        # input columns: the caller's slab (parity check against the GPU result) or seeded random data
        self.pick = [int(k) for k in pick]
        self.sa = SampledApply(mats, idx, cfg["kind"], cfg["mu"], pick, get_column or (lambda k: rng.random_sample(N)))
        self.A, self.w, self.only = self.sa.A, self.sa.w, self.sa.only
        self.n_cols = len(pick)

    def run(self):
        """(GDoF/s, seconds) for one pass over the sampled columns; the result is kept in self.result."""
        t0 = time.perf_counter()
        v = self.A.apply(self.w, only=self.only)
        dt = time.perf_counter() - t0
        self.result = {k: v[mu].coeffs for k, mu in zip(self.pick, self.only)}
        return self.N * self.n_cols / dt / 1e9, dt


# ---- GPU arm ----------------------------------------------------------------------------------------
def _bind_to_gpu_numa_node(torch, local):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is
    allocated: with several ranks the host<->device copies of the e2e leg otherwise cross the socket
    interconnect for the GPUs of the other socket.  Best effort (sysfs); returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(local)
        dev = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + dev
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus)
    torch.cuda.set_device(local)
    numa = _bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic, lambda_tables
    from spuq_b200.multi_operator import DeviceBlocks
    dev = _abi.device()
    cfg = CONFIGS[args.config]
    n, M = cfg["n"], cfg["M"]
    N = n * n
    idx = index_set(cfg)
    L = idx.shape[0]
    rvs = synthetic.make_rvs(cfg["kind"], M, cfg["mu"])
    tables = lambda_tables.LambdaTables(idx, rvs)

    rows_mode = world > 1 and args.shard == "rows"
    g = torch.Generator(device=dev); g.manual_seed(1234 + (rank if rows_mode else 0))
    if rows_mode:
        # spatial rows sharded over the ranks, one halo grid row on each side, halo exchange per step
        from spuq_b200.sharding import RowShardedApply, row_ranges
        # balanced bands, NOT rounded to the kernel's tile height: 125 rows + 2 halo rows are 16 tile rows,
        # 128 + 2 would spill into a seventeenth
        y0, y1 = row_ranges(n, world, align=1)[rank]
        nyl = y1 - y0 + 2
        Nloc = nyl * n
        rowptr, colidx, vals = synthetic.fd_blocks_rows(n, M, y0, y1, device=dev)
        blocks = DeviceBlocks(rowptr, colidx, vals, (Nloc, Nloc))
        sharded = RowShardedApply(blocks, rvs, n, y0, y1)
        sharded.halo = args.halo
        A = sharded.op
        Lp = L
        Wt = torch.rand((L, Nloc), dtype=torch.float64, device=dev, generator=g)
        # the halo row outside the global grid (first / last rank) is the homogeneous Dirichlet boundary
        if rank == 0:
            Wt.view(L, nyl, n)[:, 0, :] = 0.0
        if rank == world - 1:
            Wt.view(L, nyl, n)[:, -1, :] = 0.0
    else:
        rowptr, colidx, vals = synthetic.fd_blocks(n, M, device=dev)
        if cfg.get("perm"):
            rowptr, colidx, vals = synthetic.permute_blocks(rowptr, colidx, vals, synthetic.tile_permutation(n, cfg["perm"]))
        blocks = DeviceBlocks(rowptr, colidx, vals, (N, N))
        from spuq_b200.sharding import ShardedApply
        from spuq_b200 import _abi as _abi_mod
        sharded = ShardedApply(blocks, rvs, tables,             # world == 1: the whole operator
                               path=_abi_mod.PATH_ELL if cfg.get("path") == "ell" else _abi_mod.PATH_AUTO)
        A = sharded.op
        Lp = sharded.cols * world                               # columns padded for equal reduce-scatter shares
        Nloc = N
        Wt = torch.rand((L, N), dtype=torch.float64, device=dev, generator=g)
    w = sp.SlabMultiVector(idx, slab=Wt, _sorted=True)
    plan = A.plan_for(w)
    if args.variant or args.cols or args.rows:
        plan.tune(args.variant, args.cols, args.rows, 0)
    Yfull = torch.zeros((Lp, Nloc), dtype=torch.float64, device=dev)
    Yown = torch.empty((Lp // world, N), dtype=torch.float64, device=dev) if (world > 1 and not rows_mode) else None
    info = plan.info()

    def comm_before():
        if rows_mode:
            sharded.exchange_halos(Wt)

    def comm_after():
        if world > 1 and not rows_mode:
            dist.reduce_scatter_tensor(Yown, Yfull, op=dist.ReduceOp.SUM)

    def step():
        if rows_mode:
            sharded.apply(plan, Wt, Yfull[:L], overlap=args.overlap)   # exchange overlapped with the interior tile rows
            return
        plan.apply(Wt, Yfull[:L])
        comm_after()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for k in range(args.steps):
        kev[k][0].record()
        step()
        kev[k][1].record()
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    per_step_launches = sharded.last_launches if rows_mode else plan.last_launches
    launches = args.steps * per_step_launches

    # ---- e2e: host buffers through the public operator API, H2D + kernel + D2H every step -------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        h_in = torch.empty((L, Nloc), dtype=torch.float64, pin_memory=True)
        h_out = torch.empty((L, Nloc) if (rows_mode or world == 1) else (Lp // world, N), dtype=torch.float64, pin_memory=True)
        h_in.copy_(Wt)
        torch.cuda.synchronize()

        pipelined = (world == 1 or rows_mode)
        if pipelined:
            # H2D of step k+1 and D2H of step k-1 overlap the kernel of step k (PCIe is full duplex):
            # two device input slabs, copies on their own streams, every step still moves its whole
            # input from pinned host memory and its whole result back.
            cur = torch.cuda.current_stream()
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            Wdev = [Wt, torch.empty_like(Wt)]
            wv = [w, w.like(Wdev[1])]
            ev_in = [torch.cuda.Event(), torch.cuda.Event()]
            ev_done = [torch.cuda.Event(), torch.cuda.Event()]
            ev_out = torch.cuda.Event()
            for e in ev_done:
                e.record(cur)

            def e2e_step(k):
                b = k % 2
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_done[b])                   # the kernel that last read this slab is done
                    Wdev[b].copy_(h_in, non_blocking=True)        # H2D of this step's input slab
                    ev_in[b].record(s_in)
                cur.wait_event(ev_in[b])
                if rows_mode:
                    y = wv[b].like(torch.empty_like(Wdev[b]))
                    sharded.apply(plan, Wdev[b], y.slab, overlap=args.overlap)
                else:
                    y = A.apply(wv[b])                            # public API: MultiOperator.apply
                ev_done[b].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[b])
                    h_out.copy_(y.slab, non_blocking=True)        # D2H of the result
                    y.slab.record_stream(s_out)
                    ev_out.record(s_out)
                return y
        else:
            def e2e_step(k):
                Wt.copy_(h_in, non_blocking=True)                 # H2D of this step's input slab
                y = A.apply(w)                                    # public API: MultiOperator.apply
                Yfull[:L].copy_(y.slab)
                dist.reduce_scatter_tensor(Yown, Yfull, op=dist.ReduceOp.SUM)
                h_out.copy_(Yown, non_blocking=True)              # D2H of the rank's share of the result
                return y

        e2e_step(0); e2e_step(1)
        if pipelined:
            torch.cuda.current_stream().wait_event(ev_out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(e2e_steps):
            e2e_step(k)
        if pipelined:
            torch.cuda.current_stream().wait_event(ev_out)
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te[0]) / e2e_steps
        e2e = {"value": N * L / (e2e_ms * 1e-3) / 1e9, "unit": "GDoF/s", "steps": e2e_steps,
               "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h_in.numel() * 8),
               "d2h_bytes_per_step": int(h_out.numel() * 8)}
        launches += e2e_steps * per_step_launches

    # ---- inputs of the parity check: the few slab columns the sampled oracle apply reads, and the same
    # output columns of this run's result, gathered to rank 0 (rows mode: band by band)
    parity_in = None
    if not args.no_cpu and (world == 1 or rows_mode):
        from oracle import synth as osynth
        oidx = osynth.index_set(cfg["lam"])
        ncols = args.cpu_cols if world == 1 else min(args.cpu_cols, 4)
        pick = [int(k) for k in np.unique(np.linspace(0, L - 1, min(ncols, L)).astype(int))]
        touched = sorted(osynth.touched_columns(oidx, pick))
        step()
        torch.cuda.synchronize()

        def gather_cols(T, cols):
            if world == 1:
                return {k: T[k].cpu().numpy() for k in cols}
            band = T.view(T.shape[0], nyl, n)[cols, 1:-1, :].cpu().numpy()       # owned grid rows only
            parts = [None] * world if rank == 0 else None
            dist.gather_object(band, parts, dst=0)
            if rank != 0:
                return None
            full = np.concatenate(parts, axis=1).reshape(len(cols), -1)
            return {k: full[i] for i, k in enumerate(cols)}
        parity_in = (ncols, gather_cols(Wt, touched), gather_cols(Yfull[:L], pick))

    # ---- row-sharded PCG of config 4 on the same ranks (all ranks take part; rank 0 reports it) ----------------
    kernel_name = plan.kernel_name
    pcg_rows = None
    if rows_mode and not args.no_pcg:
        # the apply slabs go back to torch's caching allocator (NOT to the driver: the neighbours still hold
        # IPC mappings of them), the solve below allocates its own
        barrier()
        Wt = Yfull = w = plan = A = sharded = blocks = vals = Wdev = wv = h_in = h_out = y = None
        pcg_rows = pcg_block_rows(args, world, rank, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    achieved = info["algorithmic_bytes"] / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic per launch of the same kernel on the same workload, from the committed
    # `ncu --set full` capture (profiles/): dram__bytes_read.sum + dram__bytes_write.sum
    traffic, traffic_src = None, None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tr = json.load(f).get("%s|%s" % (args.config, kernel_name))
            if tr and world == 1:
                traffic = tr["dram_bytes"]
                traffic_src = ("profiles/%s: dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on "
                               "this workload from a committed `ncu --set full` capture (NOT measured in this run: a "
                               "run under ncu is never a bench value)" % name)
                break
        except Exception:
            pass
    out = {
        "metric": "SG operator apply throughput (N*|Lambda| per apply)",
        "value": N * L / (ms_per_step * 1e-3) / 1e9, "unit": "GDoF/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.config, N, L, M),
        "parallelism": ("1 GPU" if world == 1 else
                        ("grid rows sharded x%d, halo rows %s every step" % (
                            world, "written into the neighbours' slabs (CUDA IPC peer copies + flags)" if args.halo == "peer"
                            else "exchanged point-to-point (NCCL)" + (", overlapped with the interior tile rows" if args.overlap else ""))) if rows_mode
                        else ("M-sharded x%d + NCCL reduce-scatter over Lambda" % world)),
        "details": {"n_terms": info["n_terms"], "nnz_per_block": info["nnz"], "kernel": kernel_name,
                    "path": info["path"], "host_numa_node_rank0": numa},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes": info["algorithmic_bytes"], "kernel_ms": kernel_ms,
                     "fp64_gflops": info["algorithmic_flops"] / (kernel_ms * 1e-3) / 1e9},
        "clocks": clocks, "gpu_launches": launches,
    }
    if e2e:
        out["e2e"] = e2e
    # ---- parity of this run's result against the oracle on sampled output columns, and the CPU baseline ----
    # rank 0 owns the JSON line; in rows mode the touched input columns and the sampled output columns are
    # gathered band by band (only those few columns travel)
    if parity_in is not None:
        ncols, win, yout = parity_in
        ref = CpuReference(cfg, ncols, get_column=lambda k: win[k])
        v, secs = ref.run()
        worst = 0.0
        for k, r in ref.result.items():
            worst = max(worst, float(np.linalg.norm(yout[k] - r) / np.linalg.norm(r)))
        out["parity_relerr"] = worst
        out["parity_note"] = ("max over %d sampled output columns of ||y_gpu - y_oracle|| / ||y_oracle|| on this run's "
                              "input slab; oracle = oracle.sg.MultiOperator.apply(only=...)" % len(ref.result))
        if world == 1:
            out["cpu_baseline"] = {"value": v, "unit": "GDoF/s", "cores": 1, "kind": "port",
                                   "sample": "%d of %d output columns of the same workload through the oracle port of "
                                             "MultiOperator.apply (cached scipy CSR leaves, one SpMV per (mu, m)); "
                                             "%.1f s; per-column cost is independent of the other columns" % (ref.n_cols, L, secs),
                                   "host_cpus": os.cpu_count()}
    if pcg_rows is not None:
        out["pcg"] = pcg_rows
    if world == 1 and not args.no_pcg:
        # free the apply benchmark's slabs before the solve (config 4 needs ~9 slabs of 10 GB)
        # the apply slabs go back to torch's caching allocator (NOT to the driver: the neighbours still hold
        # IPC mappings of them), the solve below allocates its own
        barrier()
        Wt = Yfull = w = plan = A = sharded = blocks = vals = Wdev = wv = h_in = h_out = y = None
        out["pcg"] = pcg_block(args, peak)
    print(json.dumps(out), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def pcg_block(args, peak_gbs):
    """BASELINE.json config 4 on one GPU: PCG to 1e-10 with the exact mean-based preconditioner through
    the public API (MultiOperator / PreconditioningOperator / pcg).  Manufactured solution x* ~ U(0,1),
    f = A x*, w0 = 0.  `seconds` is the wall time of the `pcg` call (device-resident loop, the host polls
    a status word every 8 passes), `bytes_model` the compulsory traffic of one pass:
    apply + preconditioner (read rho, write s) + the fused vector updates and the two dots."""
    import torch
    import spuq_b200 as sp
    from spuq_b200 import _abi, synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    key = args.pcg_config
    cfg = CONFIGS[key]
    dev = _abi.device()
    n, M = cfg["n"], cfg["M"]
    N = n * n
    idx = index_set(cfg)
    L = idx.shape[0]
    rvs = synthetic.make_rvs(cfg["kind"], M, cfg["mu"])
    rowptr, colidx, vals = synthetic.fd_blocks(n, M, device=dev)
    A = sp.MultiOperator.from_blocks(DeviceBlocks(rowptr, colidx, vals, (N, N)), rvs)
    g = torch.Generator(device=dev); g.manual_seed(1234)
    X = torch.rand((L, N), dtype=torch.float64, device=dev, generator=g)
    x_true = sp.SlabMultiVector(idx, slab=X, _sorted=True)
    f = A.apply(x_true)
    info = A.plan_for(x_true).info()
    A0 = synthetic.blocks_to_scipy(rowptr, colidx, vals[:1], N)[0]
    P = sp.PreconditioningOperator("mean", lambda b, f_: sp.ScipySolveOperator(A0, b, b))
    w0 = 0 * f
    s = P * f                                   # set-up of the solver outside the timed solve
    del s
    hist = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, zeta, it = sp.pcg(A, f, P, w0, eps=args.pcg_eps, maxiter=200, poll_every=8, history=hist)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    solve_launches = A.plan_for(x_true).last_launches          # kernels of the whole solve (spuq_sg_pcg counts them)
    err = float((x.slab - X).abs().max())
    passes = max(it - 1, 1)

    def ms(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    t_apply = ms(lambda: A.apply(x))
    t_prec = ms(lambda: P * f)
    model = info["algorithmic_bytes"] + (16 + 88) * N * L
    return {"workload": workload_name(key), "eps": args.pcg_eps, "numit": it, "zeta": zeta, "seconds": dt,
            "ms_per_pass": dt / passes * 1e3, "gdof_per_s_per_pass": N * L * passes / dt / 1e9,
            "max_abs_error_vs_manufactured": err, "ms_apply": t_apply, "ms_precond": t_prec,
            "bytes_model_per_pass": model, "bytes_model_frac": model / (dt / passes) / 1e9 / peak_gbs,
            "preconditioner": "exact A_0^{-1} (strip partition, csrc/precond_strip.cuh)",
            # spuq_sg_pcg enqueues passes in groups of poll_every = 8 (the ones after convergence are no-ops but still
            # launches); 9 launches precede the loop: first apply, f - A w, strips forward / separator system / strips
            # final, copy, first dot, first convergence test (+ the apply's own count)
            "gpu_launches": solve_launches,
            "launches_per_pass": round((solve_launches - 9) / (-(-passes // 8) * 8), 1)}


def pcg_block_rows(args, world, rank, dev):
    """The config-4 solve with the grid rows sharded over the ranks (spuq_b200.sharding.RowShardedPCG): halo rows
    written into the neighbours' memory, scalars all-reduced in the device state, strip solver on every band with
    the separator system column-sharded.  Manufactured solution per band (seeded per rank), f = A x*, w0 = 0."""
    import torch
    import torch.distributed as dist
    import spuq_b200 as sp
    from spuq_b200 import synthetic
    from spuq_b200.multi_operator import DeviceBlocks
    from spuq_b200.sharding import RowShardedApply, RowShardedPCG, row_ranges
    key = args.pcg_config
    cfg = CONFIGS[key]
    n, M = cfg["n"], cfg["M"]
    idx = index_set(cfg)
    L = idx.shape[0]
    rvs = synthetic.make_rvs(cfg["kind"], M, cfg["mu"])
    ranges = row_ranges(n, world, align=1)
    y0, y1 = ranges[rank]
    nyl = y1 - y0 + 2
    rowptr, colidx, vals = synthetic.fd_blocks_rows(n, M, y0, y1, device=dev)
    sh = RowShardedApply(DeviceBlocks(rowptr, colidx, vals, (nyl * n, nyl * n)), rvs, n, y0, y1)
    sh.halo = args.halo
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    X = torch.zeros((L, nyl, n), dtype=torch.float64, device=dev)
    X[:, 1:-1, :] = torch.rand((L, y1 - y0, n), dtype=torch.float64, device=dev, generator=g)
    X = X.view(L, nyl * n)
    w = sp.SlabMultiVector(idx, slab=X, _sorted=True)
    plan = sh.plan_for(w)
    F = torch.empty_like(X)
    sh.apply(plan, X.clone(), F)
    solver = RowShardedPCG(sh, n, n, ranges, mean_stencil=(4.0, -1.0, -1.0))
    solver._zero_halos(F)
    tmp = torch.empty_like(F)
    solver.precondition(F, tmp)                        # set-up of the solver outside the timed solve
    del tmp
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    Xs, zeta, it = solver.solve(plan, F, torch.zeros_like(F), eps=args.pcg_eps, maxiter=200)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    err = (Xs.view(L, nyl, n)[:, 1:-1, :] - X.view(L, nyl, n)[:, 1:-1, :]).abs().max().reshape(1)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt[0])
    passes = max(it - 1, 1)
    return {"workload": workload_name(key), "eps": args.pcg_eps, "numit": it, "zeta": zeta, "seconds": dt,
            "ms_per_pass": dt / passes * 1e3, "gdof_per_s_per_pass": n * n * L * passes / dt / 1e9,
            "max_abs_error_vs_manufactured": float(err[0]),
            "sharding": "grid rows x%d; preconditioner: exact A_0^{-1}, %s" % (world, solver.mean_solver)}


def _ref_worker(ref, only, q):
    t0 = time.perf_counter()
    ref.A.apply(ref.w, only=only)
    q.put(time.perf_counter() - t0)


def run_reference(args):
    """The reference's CPU path (oracle port) on the host cores.  The reference is single-threaded; its
    outer loop over output columns (multi_operator2.py:52) is embarrassingly parallel, so the bounded
    sample of columns is spread over `--ref-procs` forked processes (copy-on-write share of the
    matrices) and the step time is the slowest process's."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cfg = CONFIGS[args.config]
    procs = max(1, min(args.ref_procs or (os.cpu_count() or 1), 32))
    ref = CpuReference(cfg, args.ref_cols * procs)
    N, L, ncols = ref.N, ref.L, ref.n_cols
    parts = [ref.only[i::procs] for i in range(procs)]
    parts = [p for p in parts if p]
    ctx = mp.get_context("fork")
    secs = []
    for k in range(args.warmup + args.steps):
        q = ctx.Queue()
        ws = [ctx.Process(target=_ref_worker, args=(ref, part, q)) for part in parts]
        t0 = time.perf_counter()
        for w in ws:
            w.start()
        for w in ws:
            w.join()
        dt = time.perf_counter() - t0
        busy = [q.get() for _ in ws]
        if k >= args.warmup:
            secs.append(max(busy) if busy else dt)
    ms = statistics.mean(secs) * 1e3 * (L / ncols)          # one full apply, extrapolated linearly in L
    value = N * L / (ms * 1e-3) / 1e9
    out = {"impl": "reference", "metric": "SG operator apply throughput (N*|Lambda| per apply)",
           "value": value, "unit": "GDoF/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": bench_config(args.config, N, L, cfg["M"]),
           "cpu_baseline": {"value": value, "unit": "GDoF/s", "cores": len(parts), "kind": "port",
                            "sample": "each step = %d of %d output columns through the oracle port of the reference's "
                                      "MultiOperator.apply (Python + scipy SpMV as the reference; its loop over output "
                                      "columns spread over %d forked processes); ms_per_step extrapolated linearly to all "
                                      "columns" % (ncols, L, len(parts)),
                            "host_cpus": os.cpu_count()},
           "e2e": {"value": value, "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), file=_JSON_OUT, flush=True)


_JSON_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--shard", default="rows", choices=["rows", "m"],
                    help="multi-GPU decomposition: grid rows + halo exchange (default) or expansion terms + reduce-scatter")
    ap.add_argument("--no-clocks", action="store_true", help="development: do not run the nvidia-smi clock sampler")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="rows mode: halo rows written into the neighbours' memory through CUDA IPC (default) or sent with NCCL")
    ap.add_argument("--overlap", action="store_true",
                    help="rows mode with --halo nccl: run the exchange beside the tile rows that read no halo row")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pcg", action="store_true", help="skip the config-4 PCG solve appended to the 1-GPU line")
    ap.add_argument("--pcg-config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--pcg-eps", type=float, default=1e-10)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-cols", type=int, default=24)
    ap.add_argument("--ref-cols", type=int, default=4, help="sampled output columns per process of the reference arm")
    ap.add_argument("--ref-procs", type=int, default=0, help="processes of the reference arm (0 = all host cpus, at most 32)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: libraries that print there (NCCL's version banner
    # under NCCL_DEBUG=VERSION) are sent to stderr for the duration of the run
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
