#!/usr/bin/env python
"""bench.py — headline benchmark of the nano_gicp scan-to-map hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): scan-to-submap GICP — one synthetic Ouster OS1-64 scan
(65,536 points, 1024x64, no voxel filter) registered against a 1,000,000-point concatenated keyframe
submap whose index and per-keyframe covariances are resident ("covariance reuse",
reference src/dlio/src/dlio/odom.cc:1719-1738). One STEP = what DLIO does per scan
(odom.cc:721-722,1005): setInputSource (K1 index build) + calculateSourceCovariances (K2 k-NN, K3
covariance) + align (K4 linearise / K5 error per LM iteration, 6x6 solve on the host).
  value : scans/s, scan already in HBM as float4 when the timed region starts (CUDA events on the
          handle's stream, L2 flushed between steps).
  e2e   : scans/s through the reference-facing API with HOST buffers (32-byte dlio::Point AoS in,
          pose out): pack + H2D + kernels + D2H inside the timed region (wall clock around the
          synchronous call).
N > 1 (torchrun, one rank per GPU): every rank registers its own sequence against its own submap
(no data-path collective) — weak scaling, value = total scans/s over the max-over-ranks time.
Beside the headline the line carries the BASELINE configs that shard (strong scaling, fixed total work):
  bulk.all_ranks       cfg 3: 256 keyframes x 65,536 points = 16,777,216 points, keyframe i -> rank i mod N
  multi_sequence_8     cfg 5: 8 seeded MulRan-shaped sequences through the odom loop, sequence i -> rank i mod N
and, on rank 0 at N = 1, cfg 4: a 500-scan OS1-64 trajectory through the odom loop over the CUDA path and over the
CPU oracle with their scan-by-scan agreement counters.
--impl reference: the CPU oracle (reference nanoflann.h + restated nano_gicp, OpenMP on all host
cores) on the same workload; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N_SCAN = 65536
N_SUBMAP = 1_000_000
N_KEYFRAMES = 40
N_DISTINCT_SCANS = 8
K_CORR = 16
CFG3_KEYFRAMES = 256          # BASELINE config 3: 256 x 65,536 = 16,777,216 points, sharded by keyframe
CFG5_SEQUENCES = 8            # BASELINE config 5
CFG5_SCANS = int(os.environ.get("NGICP_BENCH_CFG5_SCANS", 200))
CFG4_SCANS = int(os.environ.get("NGICP_BENCH_CFG4_SCANS", 500))
# algorithmic bytes per unit: compulsory traffic only
BYTES = {
    "K1_index_per_pt": 36,        # 16 R + 16 W reordered float4 + 4 W permutation
    "K2_knn_per_pt": 16 + 4 * K_CORR + 8,   # 16 R + k*4 W neighbour ids + 8 W density term (distances are not materialised)
    "K3_cov_per_pt": 16 + 4 * K_CORR + 24,  # 104
    "K4a_corr_per_src_pt": 20,    # correspondence search: 16 R p_A + 4 W correspondence (the target is gathered, counted 0)
    "K4b_lin_per_src_pt": 84,     # fused linearisation: 16 p_A + 24 C_A + 4 corr + 16 p_B + 24 C_B
    "K5_err_per_src_pt": 84,      # 16 p_A + 4 corr + 24 C_A + 16 p_B + 24 C_B (Mahalanobis rebuilt, not cached)
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner there) are sent to stderr
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# ------------------------------------------------------------------------------------------- data
def make_workload(seed: int):
    """(submap (1M,3), keyframe bounds, list of 8 perturbed world-frame scans (65536,3))."""
    from ngicp import synth
    sc = synth.Scene(seed)
    rng = np.random.default_rng(seed + 2)
    tgt, bounds, poses = synth.make_submap(sc, N_SUBMAP, seed, n_keyframes=N_KEYFRAMES)
    scans = []
    for i in range(N_DISTINCT_SCANS):
        T_ws = poses[(5 + 4 * i) % len(poses)] @ synth.se3((0, 0, 0.02), (0.3, 0.1, 0.0))
        s = synth.transform_points(T_ws, synth.scan(sc, T_ws, rng, keep_all=True))
        T_off = synth.random_se3(rng, 0.2, 2.0)        # DLIO hands GICP a scan already close to the map (odom.cc:1005)
        scans.append(synth.transform_points(np.linalg.inv(T_off), s))
    return tgt, bounds, scans


def ncu_traffic(name):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            v = json.load(f).get(name)
        return None if v is None else float(v)
    except (OSError, ValueError):
        return None


def configure(g):
    import scenarios as S
    return S.configure(g, k=K_CORR, max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01)   # cfg/params.yaml:57-63


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md; its period, 200 ms). Every query
    takes driver-wide locks: at -lms 10 the sampler of rank 0 stalled kernel launches of the OTHER ranks for 8-12 ms at a
    time (measured at N = 8: the cfg-3 index build 0.72 ms on some ranks, 8-12 ms on others; gone without the sampler)."""

    def __init__(self, index: int, cores=None):
        self.index = index
        self.rows = []
        self.proc = None
        self.cores = set(cores) if cores else None     # where nvidia-smi and the reader thread may run (not the launching thread's core)

    def _move(self):
        if self.cores:
            try:
                os.sched_setaffinity(0, self.cores)
            except OSError:
                pass

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, preexec_fn=self._move)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:   # nvidia-smi needs a moment to come up
                time.sleep(0.01)
            self.skip = len(self.rows)                        # samples before the timed region are dropped
        except Exception:
            self.proc = None

    def _read(self):
        self._move()
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.06)                                  # let the sample covering the end of the region land
            self.proc.terminate()
        self.rows = self.rows[getattr(self, "skip", 0):] or self.rows[-1:]
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arms
def cpu_register_stream(tgt, bounds, scans, steps, warmup):
    """The reference's CPU path on the same workload. Returns (scans/s, description)."""
    import oracle
    variant = "ref" if oracle.available("ref") else "port"
    threads = os.cpu_count() or 1          # all host cores, whatever OMP_NUM_THREADS torchrun put in the environment
    o = configure(oracle.OracleGICP(variant, num_threads=threads))
    # submap: per-keyframe covariances computed once and concatenated (covariance reuse), then the tree
    covs = []
    for s, e in zip(bounds[:-1], bounds[1:]):
        o.setInputSource(tgt[s:e]); o.calculateSourceCovariances(); covs.append(o.getSourceCovariances())
    o.setInputTarget(tgt); o.setTargetCovariances(np.concatenate(covs))
    ts = []
    for i in range(warmup + steps):
        src = scans[i % len(scans)]
        t0 = time.perf_counter()
        o.setInputSource(src); o.calculateSourceCovariances(); o.align()
        ts.append(time.perf_counter() - t0)
    ts = ts[warmup:]
    kind = "port"   # nano_gicp.cc cannot be compiled here (Eigen/PCL absent); the k-NN inside IS the reference's nanoflann when variant == ref
    desc = (f"{steps} scan registrations (65,536-pt scan vs 1,000,000-pt submap, k=16) with the oracle "
            f"({'reference nanoflann.h KD-tree' if variant == 'ref' else 'port k-d tree'} + restated nano_gicp, OpenMP guided,8)")
    return steps / sum(ts), 1e3 * sum(ts) / steps, threads, kind, desc


def run_reference(args, rank):
    if rank != 0:
        return
    tgt, bounds, scans = make_workload(0)
    val, ms, threads, kind, desc = cpu_register_stream(tgt, bounds, scans, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "gicp_scan_to_submap_scans_per_s", "value": val, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(1),
            "cpu_baseline": {"value": val, "unit": "scans/s", "cores": threads, "kind": kind, "sample": desc},
            "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(world):
    return {"workload": "cfg2 scan-to-submap GICP: 65,536-pt synthetic OS1-64 scan vs 1,000,000-pt keyframe submap (40 keyframes, covariance reuse), "
                        "k=16, max_corr 0.5 m, max_iter 32, eps 0.01/0.01; step = setInputSource + calculateSourceCovariances + align",
            "sequences": world, "parallelism": f"independent sequences x{world} (no collective); every rank registers the same seeded scans against its own resident copy of the submap",
            "l2": "flushed between timed steps (256 MiB write)"}


# ------------------------------------------------------------------------------------------- GPU arm
def physical_cores(cpus):
    """The allowed logical CPUs grouped by physical core (hyper-thread siblings together), from sysfs."""
    groups, seen = [], set()
    for c in sorted(cpus):
        if c in seen:
            continue
        sib = {c}
        try:
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as f:
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    sib.update(range(int(a), int(b or a) + 1))
        except (OSError, ValueError):
            pass
        sib &= set(cpus)
        seen |= sib
        groups.append(sorted(sib))
    return groups


_RANK_CORES = {"groups": None}


def pin_rank_to_cores(local_rank, world):
    """Disjoint PHYSICAL cores per rank: the LM loop is ~10 launch -> poll round trips per scan, and eight unpinned processes
    on the same cores — or two polling threads on the two hyper-threads of one core — turn into step-time loss at N > 1."""
    try:
        groups = physical_cores(os.sched_getaffinity(0))
        per = max(1, len(groups) // max(world, 1))
        mine = groups[local_rank * per:(local_rank + 1) * per] or groups
        _RANK_CORES["groups"] = mine
        cpus = sorted(c for g in mine for c in g)
        os.sched_setaffinity(0, cpus)
        return cpus
    except (AttributeError, OSError):
        return None


def dedicate_core_to_main_thread(cores):
    """The launching thread gets one physical core of the rank's slice to itself (the last one: core 0 of a box takes its
    interrupts and housekeeping); every helper thread that exists by now (CUDA, NCCL proxy / watchdog, BLAS pools) is moved
    to the other cores, the hyper-thread sibling of the launching thread's CPU stays idle. A 2-3 ms preemption of the
    launch -> poll loop is a quarter of a 20-step timed region."""
    try:
        groups = _RANK_CORES["groups"]
        if not groups or len(groups) < 2:
            return None
        main = threading.get_native_id()
        rest = {c for g in groups[:-1] for c in g}
        for t in os.listdir("/proc/self/task"):
            tid = int(t)
            try:
                os.sched_setaffinity(tid, {groups[-1][0]} if tid == main else rest)
            except OSError:
                pass
        try:
            os.setpriority(os.PRIO_PROCESS, main, -20)      # the spinning launch thread should not lose its core to a waking daemon
        except (OSError, AttributeError):
            pass
        return sorted(rest)
    except (AttributeError, OSError):
        return None


def dropin_cpp_e2e(tgt, bounds, scans, steps, warmup):
    """The C++ drop-in surface (include/nano_gicp/nano_gicp.h) driven the way dlio::OdomNode drives the reference, on fresh
    pageable pcl::PointCloud scans: tests/shims/dropin_bench.cc compiled here against the PCL/Eigen shims (PCL is absent)."""
    import tempfile
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    lib = ROOT / "noetic-slam_b200" / "libngicp_b200.so"
    out = {}
    with tempfile.TemporaryDirectory() as td:
        exe = Path(td) / "dropin_bench"
        cmd = [cxx, "-std=c++17", "-O2", "-DSHIM_PCL_MINOR=10", f"-I{ROOT / 'tests' / 'shims'}", f"-I{ROOT / 'include'}",
               str(ROOT / "tests" / "shims" / "dropin_bench.cc"), "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            return {"error": "compile failed: " + r.stderr[-300:]}
        np.ascontiguousarray(tgt, np.float32).tofile(Path(td) / "tgt.bin")
        np.ascontiguousarray(bounds, np.int64).tofile(Path(td) / "bounds.bin")
        names = []
        for i, sc in enumerate(scans):
            np.ascontiguousarray(sc, np.float32).tofile(Path(td) / f"scan{i}.bin")
            names.append(str(Path(td) / f"scan{i}.bin"))
        for key, flag in (("align_fills_output_cloud", "1"), ("output_cloud_skipped", "0")):
            r = subprocess.run([str(exe), str(Path(td) / "tgt.bin"), str(Path(td) / "bounds.bin"), str(steps), str(warmup), flag] + names,
                               capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                out[key] = {"error": (r.stderr or r.stdout)[-300:]}
                continue
            out[key] = json.loads(r.stdout.strip().splitlines()[-1])
    return out


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import ngicp
    from ngicp import sharding, synth

    cores = pin_rank_to_cores(local_rank, world)
    # ---- host data first: the sequence generators fork worker processes, which must happen before CUDA is initialised
    t_gen = time.time()
    # every rank gets the SAME seeded workload (its own resident copy of the submap, the same eight scans in the same order):
    # per-GPU work is then exactly fixed as N grows, and the sharded legs (cfg 3 keyframes, cfg 5 sequences) partition one
    # global job that is the same at every N. (With per-rank seeds the max-over-ranks time is set by the hardest sequence:
    # 0.92 of linear at N = 8 from heterogeneity alone.)
    tgt, bounds, scans = make_workload(0)
    owned_seq = sharding.units_for_rank(CFG5_SEQUENCES, rank, world)
    specs = [(100 + k, CFG5_SCANS, 0.4, 1, True) for k in owned_seq]          # cfg 5: MulRan-shaped (t = 0), one deskew group
    with_cfg4 = world == 1 and CFG4_SCANS > 0
    if with_cfg4:
        specs.append((4, CFG4_SCANS, 0.4, 2, False))                           # cfg 4: OS1-64 time stamps, sensor moving during the scan
    seqs = generate_sequences(specs, world) if (CFG5_SCANS > 0 or with_cfg4) else []
    seq_cfg4 = seqs.pop() if with_cfg4 else None
    log(f"[rank {rank}] workload + {len(specs)} sequences generated in {time.time() - t_gen:.1f}s (cores {cores[:1]}..{cores[-1:] if cores else ''})")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    g = configure(ngicp.NanoGICP(local_rank))
    # submap resident in HBM: per-keyframe covariances in ONE batched pass (reuse), then the 1M-pt index
    t0 = time.time()
    _, m4, _ = g.batchCovariances(tgt, bounds, want_mat4=True)
    g.setInputTarget(tgt)
    g.setTargetCovariances(m4)
    g.synchronize()
    log(f"[rank {rank}] submap index + keyframe covariances resident in {time.time() - t0:.2f}s")

    stream = torch.cuda.ExternalStream(g.stream_ptr(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    d_scans = []
    for s in scans:
        f4 = np.zeros((len(s), 4), np.float32)
        f4[:, :3] = s
        d_scans.append(torch.from_numpy(f4).to(dev))
    h_scans = [synth.to_aos32(s) for s in scans]       # the reference's 32-byte AoS
    pinned = [torch.empty(h_scans[0].shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    torch.cuda.synchronize()

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_job(t, u):
        return sharding.reduce_job(t, u, device=dev)

    W, K = max(3, args.warmup), args.steps     # never fewer than three warm-up steps (timing rules)
    # ---- value: device-resident input, CUDA events on the handle's stream
    sampler = ClockSampler(local_rank, [c for g in (_RANK_CORES["groups"] or [])[:-1] for c in g] or None) if rank == 0 and os.environ.get("NGICP_BENCH_SAMPLER", "1") != "0" else None      # one sampler per job, not one per rank
    iters = []
    g.timings(reset=True)
    launches0 = 0
    dts = []
    if world > 1:
        # NCCL sets its connections up lazily, inside the first collectives, and tidies up (buffer registration, proxy
        # threads, device allocations and frees — the latter synchronise this rank's GPU) for a few milliseconds afterwards:
        # take that out of the way now, so that the barrier in front of the timed region is a steady-state collective
        for _ in range(3):
            dist.barrier()
            warm_t = torch.zeros(8, dtype=torch.float64, device=dev)
            dist.all_reduce(warm_t)
            dist.all_reduce(warm_t, op=dist.ReduceOp.MAX)
            torch.cuda.synchronize()
    if os.environ.get("NGICP_BENCH_DEDICATE", "1") != "0":
        dedicate_core_to_main_thread(cores)
    for i in range(W + K):
        if i == W:
            if sampler:
                sampler.start()
            barrier()
            launches0 = g.timings(reset=False)["kernel_launches"]
        flush_l2()
        ds = d_scans[i % len(d_scans)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        g.setInputSourceDevice(ds.data_ptr(), ds.shape[0], token=ds)
        g.calculateSourceCovariances()
        g.align()
        e1.record(stream)
        e1.synchronize()
        dts.append(e0.elapsed_time(e1) * 1e-3)
        iters.append(g.nr_iterations_ + 1)
    barrier()
    launches = g.timings(reset=False)["kernel_launches"] - launches0
    t_local = float(sum(dts[W:]))
    t_max, units = reduce_job(t_local, float(K))
    value = units / t_max
    # per-rank spread of the step time (separates jitter from a systematic slow-down at N > 1)
    step_ms = [1e3 * float(np.median(dts[W:])), 1e3 * float(np.min(dts[W:])), 1e3 * float(np.max(dts[W:])), float(np.argmax(dts[W:]))]
    if world > 1:
        allv = torch.zeros(world, 4, dtype=torch.float64, device=dev)
        allv[rank] = torch.tensor(step_ms, dtype=torch.float64)
        dist.all_reduce(allv)
        per_rank = allv.cpu().numpy().tolist()
    else:
        per_rank = [step_ms]

    # ---- e2e: host buffers through the reference-facing API, wall clock. Two kinds of scan buffer:
    #      page-locked (a registered cloud: copied as it is) and pageable (what an unmodified DLIO hands over: a fresh
    #      pcl::PointCloud per scan, packed into the handle's pinned staging buffer)
    def e2e_loop(kind):
        ts = []
        for i in range(W + K):
            if i == W:
                barrier()
            flush_l2()
            if kind == "pinned":
                hs = pinned[i % len(pinned)].numpy()       # refilled outside the timed region; align() has synchronised its last use
                hs[:] = h_scans[i % len(h_scans)]
            else:
                hs = h_scans[i % len(h_scans)].copy()      # a fresh pageable cloud, just written by the CPU (as a deskewed scan is)
            t0 = time.perf_counter()
            g.setInputSource(hs)
            g.calculateSourceCovariances()
            g.align()
            ts.append(time.perf_counter() - t0)
        barrier()
        e_max, e_units = reduce_job(float(sum(ts[W:])), float(K))
        return e_units / e_max, 1e3 * e_max / K

    e2e_value, e2e_ms = e2e_loop("pinned")
    e2e_pg_value, e2e_pg_ms = e2e_loop("pageable")
    if cores:
        try:
            os.sched_setaffinity(0, cores)     # the legs below start their own threads: back to the rank's whole slice
            os.setpriority(os.PRIO_PROCESS, threading.get_native_id(), 0)
        except (OSError, AttributeError):
            pass

    # ---- BASELINE config 3 across the ranks (strong scaling): 256 keyframes x 65,536 points in total, keyframe i on rank i mod N
    hbm = peak_hbm()
    owned_kf = sharding.units_for_rank(CFG3_KEYFRAMES, rank, world)
    bulk = bulk_covariance_cfg3(g, scans, hbm, owned_kf)
    shard_ms = bulk["index_ms"] + bulk["knn_ms"] + bulk["covariance_ms"]
    b_t, b_pts = reduce_job(shard_ms * 1e-3, float(bulk["points"]))
    clocks = sampler.stop() if sampler else None     # sampled across the timed regions (device-resident loop, host-buffer loops, cfg-3 build)
    k3_t, _ = reduce_job(bulk["covariance_ms"] * 1e-3, 0.0)
    k2_t, _ = reduce_job(bulk["knn_ms"] * 1e-3, 0.0)
    shard = [bulk["index_ms"], bulk["knn_ms"], bulk["covariance_ms"]]
    if world > 1:
        allb = torch.zeros(world, 3, dtype=torch.float64, device=dev)
        allb[rank] = torch.tensor(shard, dtype=torch.float64)
        dist.all_reduce(allb)
        shard_per_rank = allb.cpu().numpy().round(4).tolist()
    else:
        shard_per_rank = [[round(v, 4) for v in shard]]
    bulk["all_ranks"] = {"per_rank_index_knn_cov_ms": shard_per_rank, "ranks": world, "keyframes": CFG3_KEYFRAMES, "points": int(b_pts), "covariance_build_mpts_s": b_pts / b_t / 1e6,
                         "build_ms_max_over_ranks": 1e3 * b_t, "knn_ms_max_over_ranks": 1e3 * k2_t, "K3_only_gpts_s": b_pts / k3_t / 1e9,
                         "scaling": "strong (16,777,216 points in total, keyframe i -> rank i mod N)"}

    # ---- BASELINE config 5 across the ranks (strong scaling): 8 sequences through the odom loop
    multi8 = multi_sequence_cfg5(local_rank, rank, world, seqs, owned_seq, barrier, reduce_job) if CFG5_SCANS > 0 else None
    del seqs

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel device times (instrumented pass; not part of the headline numbers)
    g.enableTiming(True)
    g.timings(reset=True)
    reps = 5
    for i in range(reps):
        flush_l2()
        ds = d_scans[i % len(d_scans)]
        g.setInputSourceDevice(ds.data_ptr(), ds.shape[0], token=ds)
        g.calculateSourceCovariances()
        g.align()
    t = g.timings(reset=True)
    g.enableTiming(False)
    per = {
        "K1_index": (t["index_ms"] / reps, BYTES["K1_index_per_pt"] * N_SCAN),
        "K2_knn": (t["knn_ms"] / reps, BYTES["K2_knn_per_pt"] * N_SCAN),
        "K3_covariance": (t["covariance_ms"] / reps, BYTES["K3_cov_per_pt"] * N_SCAN),
        "K4a_correspond": (t["correspond_ms"] / max(t["linearize_calls"], 1), BYTES["K4a_corr_per_src_pt"] * N_SCAN),
        "K4b_linearize": (t["linearize_ms"] / max(t["linearize_calls"], 1), BYTES["K4b_lin_per_src_pt"] * N_SCAN),
        "K5_error": (t["error_ms"] / max(t["error_calls"], 1), BYTES["K5_err_per_src_pt"] * N_SCAN),
    }
    step_share = {"K1_index": t["index_ms"], "K2_knn": t["knn_ms"], "K3_covariance": t["covariance_ms"], "K4a_correspond": t["correspond_ms"],
                  "K4b_linearize": t["linearize_ms"], "K5_error": t["error_ms"]}
    dominant = max(step_share, key=step_share.get)
    kernels = {k: {"ms_per_launch": ms, "algorithmic_bytes": b, "GBps": b / (ms * 1e-3) / 1e9 if ms > 0 else None,
                   "share_of_step": step_share[k] / max(sum(step_share.values()), 1e-9), "ncu_dram_bytes": ncu_traffic(k)} for k, (ms, b) in per.items()}
    dm, db = per[dominant]
    step_dominant = {"kernel": dominant, "bound": "hbm", "achieved": db / (dm * 1e-3) / 1e9, "peak": hbm["gbs"], "unit": "GB/s",
                     "frac": db / (dm * 1e-3) / 1e9 / hbm["gbs"], "traffic": ncu_traffic(dominant), "peak_source": hbm["source"],
                     "note": "single-scan launches are L2/latency-bound (5 MB per launch); the HBM bar applies to the bulk numbers"}

    # ---- judged bulk numbers: K3 on the cfg-3 launch above and K4b (fused linearisation) on a batch of 64 scans against the
    #      resident submap
    bulk.update(bulk_linearize(g, scans, hbm))
    bulk.update(prefilter_probe(g, pinned, h_scans, world == 1))
    odom4 = None
    cpp = None
    if world == 1:
        bulk.update(multi_sequence_probe(g, tgt, m4, d_scans, local_rank))
        if seq_cfg4 is not None:
            odom4 = odom_loop_cfg4(local_rank, seq_cfg4)
        cpp = dropin_cpp_e2e(tgt, bounds, scans, K, W)

    # ---- CPU baseline beside it (bounded sample, all host cores) — rank 0 at N=1 only
    cpu = None
    if world == 1:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))     # the oracle gets every host core
        except (AttributeError, OSError):
            pass
        cpu_val, cpu_ms, threads, kind, desc = cpu_register_stream(tgt, bounds, scans, 10, 2)
        cpu = {"value": cpu_val, "unit": "scans/s", "ms_per_step": cpu_ms, "cores": threads, "kind": kind, "sample": desc}

    # Top-level roofline = the covariance kernel (K3) on the cfg-3 launch: the north star puts the >= 50 % HBM bar on the
    # covariance and linearisation kernels, and SURVEY.md §8(d) takes K2 (the kernel with the largest share of a single
    # step, latency/issue-bound exact k-NN) out of the HBM bar. Both judged kernels and the step's dominant kernel are listed.
    roofline = dict(bulk["roofline_K3"])
    roofline["kernel"] = "K3_covariance, cfg-3 launch (%d keyframes x %d points on this GPU)" % (bulk["keyframes"], N_SCAN)
    roofline["ms_per_launch"] = bulk["covariance_ms"]
    roofline["judged"] = {"K3_covariance_bulk": bulk["roofline_K3"], "K4b_linearize_batched": bulk["roofline_K4b"]}
    roofline["step_dominant"] = step_dominant
    mean_it = float(np.mean(iters[W:]))
    d2h = int(4 + mean_it * (29 * 8 + 8) + 2.5 * mean_it * 16 + 64)   # density + per LM iteration one 29-double result row (+ seq) and ~2.5 error rows + pose
    line = {
        "metric": "gicp_scan_to_submap_scans_per_s", "value": value, "unit": "scans/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world),
        "ms_per_align_step": 1e3 * t_max / K, "lm_iterations_per_scan": mean_it,
        "per_rank_step_ms_median_min_max_argmax": per_rank,
        "e2e": {"value": e2e_value, "unit": "scans/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": N_SCAN * 32,
                "host_buffer": "page-locked 32-byte AoS scan, copied as it is (cudaMemcpyAsync); the call returns when the copy has landed",
                "d2h_bytes_per_step": d2h, "d2h_note": "host-mapped result rows the kernels write (counted from the row sizes, not metered)"},
        "e2e_pageable": {"value": e2e_pg_value, "unit": "scans/s", "ms_per_step": e2e_pg_ms, "h2d_bytes_per_step": N_SCAN * 12,
                         "host_buffer": "fresh pageable 32-byte AoS scan per step (what an unmodified DLIO hands over): xyz packed into the pinned staging "
                                        "buffer in four slices, each slice's copy overlapping the next slice's packing"},
        "e2e_cpp_dropin": cpp,
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "kernels": kernels, "bulk": bulk, "multi_sequence_8": multi8, "odom_loop_cfg4": odom4,
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- sequences (cfg 4 / cfg 5)
_SEQ_JOBS = {}   # (seed, step) -> poses; filled before the pool forks


def _gen_scan(job):
    seed, i, step, groups, mulran = job
    from ngicp import odom, synth
    return odom.synthetic_scan(synth.Scene(seed), _SEQ_JOBS[(seed, step)], i, seed, 1024, groups, mulran)


def generate_sequences(specs, world):
    """specs: [(seed, n_scans, step_m, deskew_groups, mulran)] -> [list of synthetic_scan tuples]. Ray casting is numpy on
    the host (0.2-0.5 s per 65,536-point scan), so the scans are made by a fork pool — BEFORE this process touches CUDA."""
    import multiprocessing as mp
    from ngicp import odom, synth
    jobs = []
    for seed, n, step, groups, mulran in specs:
        _SEQ_JOBS[(seed, step)] = odom.synthetic_poses(synth.Scene(seed), n, seed, step)
        jobs += [(seed, i, step, groups, mulran) for i in range(n)]
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    procs = max(1, min(32, cores // max(world, 1)))
    if procs > 1 and len(jobs) > 8:
        with mp.get_context("fork").Pool(procs) as pool:
            flat = pool.map(_gen_scan, jobs, chunksize=4)
    else:
        flat = [_gen_scan(j) for j in jobs]
    out, at = [], 0
    for seed, n, step, groups, mulran in specs:
        out.append(flat[at:at + n]); at += n
    return out


def drive_loop(loop, seq, drift, groups, i0, i1, ts=None, res=None):
    """Scans [i0, i1) of one sequence through an OdomLoop (the prior the IMU integration would supply = the generating pose
    of every deskew group, disturbed by a seeded drift)."""
    for i in range(i0, i1):
        rec, Ts, block, col_t = seq[i]
        t0 = time.perf_counter()
        if i == 0:
            if hasattr(loop, "set_pose"):
                loop.set_pose(Ts[groups // 2])
            else:
                loop.T = Ts[groups // 2].astype(np.float32)
                loop.propagateGICP()
            r = loop.callbackPointCloud(rec, None)
        else:
            def prior(stamps, Ts=Ts, i=i):
                k = np.minimum((stamps.astype(np.int64) * groups) // 100_000_000, groups - 1)
                return (drift[i] @ Ts)[k].astype(np.float32)
            r = loop.callbackPointCloud(rec, prior)
        if ts is not None:
            ts.append(time.perf_counter() - t0)
        if res is not None:
            res.append(r)


def multi_sequence_cfg5(device, rank, world, seqs, owned, barrier, reduce_job):
    """The cfg-5 leg, timed on its second pass: an untimed pass over the first 60 scans of every sequence (own loops, own
    handles, released afterwards) first takes whatever is paid once per process — kernels that only the keyframe / submap
    paths launch, memory-pool blocks of the submap sizes — out of a timed region that is only 0.25 s long (without it the
    same leg read 2,300-6,200 scans/s from run to run; repeated in one process it reads 5,400-6,300 every time)."""
    if seqs and len(seqs[0]) > 80:
        import gc
        _multi_sequence_pass(device, rank, world, [s[:60] for s in seqs], owned, lambda: None, lambda t, u: (t, u))
        gc.collect()
    out = _multi_sequence_pass(device, rank, world, seqs, owned, barrier, reduce_job)
    out["warm_up"] = "one untimed pass over the first 60 scans of every sequence on separate loops, then this pass on fresh loops"
    return out


def _multi_sequence_pass(device, rank, world, seqs, owned, barrier, reduce_job):
    """BASELINE config 5: 8 independent MulRan-shaped sequences (t = 0: one deskew group, reference
    src/file_player_mulran/src/ROSThread.cpp:509-518) through the per-scan loop (src/dlio/src/dlio/odom.cc:737-837), sequence i on
    rank i mod N, the sequences of a rank side by side on its GPU (one handle, stream and host thread each). Strong scaling:
    the job is the same 8 sequences at every N. Wall clock between barriers (the loop is host-driven), max over ranks."""
    import ngicp
    from ngicp import odom, synth
    warm = 3
    loops, drifts = [], []
    for k, seq in zip(owned, seqs):
        g = configure(ngicp.NanoGICP(device))
        loops.append(odom.NativeOdomLoop(g, odom.OdomParams()))
        rng = np.random.default_rng(1000 + k)
        drifts.append([synth.random_se3(rng, 0.03, 0.3) for _ in range(len(seq))])
    results = [[] for _ in loops]
    for loop, seq, dr, res in zip(loops, seqs, drifts, results):      # untimed: first scans (allocations, first keyframe, first submap)
        drive_loop(loop, seq, dr, 1, 0, warm, res=res)
    gate = threading.Barrier(len(loops) + 1)

    def worker(loop, seq, dr, res):
        gate.wait()
        drive_loop(loop, seq, dr, 1, warm, len(seq), res=res)
        loop.gicp.synchronize()

    ths = [threading.Thread(target=worker, args=a) for a in zip(loops, seqs, drifts, results)]
    for t in ths:
        t.start()
    barrier()
    gate.wait()
    t0 = time.perf_counter()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    units = float(sum(len(s) - warm for s in seqs))
    t_max, u_all = reduce_job(dt, units)
    err = max((float(np.abs(r.T[:3, 3] - s[i][1][0][:3, 3]).max()) for s, res in zip(seqs, results) for i, r in enumerate(res) if r is not None and i > 0), default=0.0)
    its = [r.iterations + 1 for res in results for r in res[1:] if r is not None]
    e_max, _ = reduce_job(err, 0.0)
    return {"sequences": CFG5_SEQUENCES, "scans_per_sequence": len(seqs[0]) if seqs else 0, "timed_scans_all_ranks": int(u_all), "ranks": world,
            "sequences_on_this_rank": len(seqs), "scans_per_s": u_all / t_max, "wall_s_max_over_ranks": t_max, "scaling": "strong (8 sequences in total)",
            "keyframes_rank0": [l.n_keyframes for l in loops], "lm_iterations_mean_rank0": float(np.mean(its)) if its else None,
            "max_abs_position_error_m": e_max,
            "note": "MulRan-shaped 65,536-point records (t = 0), full per-scan loop (ingest, deskew, VoxelGrid, index, covariances, align, keyframes, "
                    "submap rebuilds); host policy in C++ (ngicp_odom_*, csrc/odom_loop.cu), the prior poses come from a Python callback; H2D of the raw records included"}


def odom_loop_cfg4(device, seq, groups=2):
    """BASELINE config 4 (bounded to CFG4_SCANS scans): a seeded OS1-64 trajectory at 10 Hz, the sensor moving during every scan,
    through the odom loop over the CUDA path and — the checker and the CPU baseline of this config — over the oracle, with
    their scan-by-scan agreement (keyframe decisions, submap sets, LM iteration counts, poses)."""
    import ngicp
    import oracle
    from ngicp import odom, synth
    from odom_backends import OracleBackend
    import scipy.spatial  # noqa: F401  (one-off import kept out of the per-scan times)
    n = len(seq)
    rng = np.random.default_rng(8)
    drift = [synth.random_se3(rng, 0.03, 0.3) for _ in range(n)]

    stages = []

    def run(backend):
        loop = odom.OdomLoop(backend, odom.OdomParams()) if backend is not None else odom.NativeOdomLoop(configure(ngicp.NanoGICP(device)), odom.OdomParams())
        ts, res = [], []
        if backend is None:           # scan by scan, to keep the stage times of every scan
            for i in range(n):
                drive_loop(loop, seq, drift, groups, i, i + 1, ts, res)
                stages.append(loop.profile(reset=True))
        else:
            drive_loop(loop, seq, drift, groups, 0, n, ts, res)
        return loop, res, ts

    lg, rg, tg = run(odom.DeviceBackend(configure(ngicp.NanoGICP(device))))
    ln, rn, tn = run(None)
    both = [(a, b) for a, b in zip(rn, rg) if a is not None and b is not None]
    cpp = {"ms_per_scan_median": 1e3 * float(np.median(tn[3:])), "ms_per_scan_mean": 1e3 * float(np.mean(tn[3:])),
           "ms_per_scan_p99": 1e3 * float(np.percentile(tn[3:], 99)), "ms_per_scan_max": 1e3 * float(np.max(tn[3:])), "scans_per_s": float(len(tn[3:]) / np.sum(tn[3:])),
           "keyframes": ln.n_keyframes,
           "stage_ms_median": {k: float(np.median([st[k] for st in stages[3:] if st["scans"]])) for k in odom.NativeOdomLoop.STAGES},
           "slowest_scans": [{"scan": int(i), "ms": 1e3 * tn[i], "new_keyframe": bool(rn[i] and rn[i].new_keyframe),
                              "stage_ms": {k: round(stages[i][k], 3) for k in odom.NativeOdomLoop.STAGES}} for i in np.argsort(-np.array(tn))[:3] if i >= 3],
           "scans_compared_with_python_loop": len(both),
           "same_decisions_as_python_loop": int(sum(a.new_keyframe == b.new_keyframe and a.submap == b.submap and a.iterations == b.iterations
                                                    and a.converged == b.converged and a.n_points == b.n_points for a, b in both)),
           "max_pose_diff_m": max(float(np.abs(a.T[:3, 3] - b.T[:3, 3]).max()) for a, b in both),
           "note": "the same loop in C++ behind the C ABI (ngicp_odom_scan_begin / _finish); the prior poses still come from a Python callback"}
    threads = os.cpu_count() or 1
    variant = "ref" if oracle.available("ref") else "port"
    lo, ro, to = run(OracleBackend(configure(oracle.OracleGICP(variant, num_threads=threads))))
    pairs = [(a, b) for a, b in zip(rg, ro) if a is not None and b is not None]
    steady = tg[3:]
    return {"scans": n, "points_per_scan": int(len(seq[0][0])), "path_m": float(np.linalg.norm(np.diff([s[1][0][:3, 3] for s in seq], axis=0), axis=1).sum()),
            "ms_per_scan_median": 1e3 * float(np.median(steady)), "ms_per_scan_mean": 1e3 * float(np.mean(steady)),
            "ms_per_scan_p99": 1e3 * float(np.percentile(steady, 99)), "ms_per_scan_max": 1e3 * float(np.max(steady)), "scans_per_s": float(len(steady) / np.sum(steady)),
            "keyframes": len(lg.keyframes), "max_submap_keyframes": max(len(r.submap) for r in rg if r is not None),
            "submap_rebuilds": int(sum(1 for r in rg if r is not None and r.submap_changed)),
            "lm_iterations_mean": float(np.mean([r.iterations + 1 for r in rg[1:] if r is not None])),
            "max_abs_position_error_m": max(float(np.abs(r.T[:3, 3] - s[1][groups // 2][:3, 3]).max()) for r, s in zip(rg[1:], seq[1:]) if r is not None),
            "cpp_loop": cpp,
            "oracle": {"kind": "port", "knn": "reference nanoflann.h" if variant == "ref" else "port k-d tree", "cores": threads,
                       "ms_per_scan_median": 1e3 * float(np.median(to[3:])), "keyframes": len(lo.keyframes),
                       "same_keyframe_decisions": int(sum(a.new_keyframe == b.new_keyframe for a, b in pairs)),
                       "same_submap_sets": int(sum(a.submap == b.submap for a, b in pairs)),
                       "same_iterations_and_convergence": int(sum(a.iterations == b.iterations and a.converged == b.converged for a, b in pairs)),
                       "scans_compared": len(pairs),
                       "max_pose_diff_m": max(float(np.abs(a.T[:3, 3] - b.T[:3, 3]).max()) for a, b in pairs),
                       "max_rotation_entry_diff": max(float(np.abs(a.T[:3, :3] - b.T[:3, :3]).max()) for a, b in pairs)},
            "note": "wall clock per callbackPointCloud incl. host policy (Python), H2D of the raw 32-byte records, D2H of the deskewed cloud"}


def bulk_covariance_cfg3(g, scans, hbm, owned):
    """BASELINE config 3: the covariances of a 16,777,216-point submap (256 keyframes x 65,536 points), keyframe i built by
    rank i mod N — this rank's keyframes in ONE batched pass (every keyframe its own index, neighbours never cross keyframes:
    the reference computes covariances per scan and concatenates them per submap, nano_gicp.cc:174-181, odom.cc:1719-1729)."""
    from ngicp import synth
    clouds = []
    for i in owned:
        rng = np.random.default_rng(9900 + i)          # keyframe i is the same cloud whatever the number of ranks
        T = synth.se3((0, 0, rng.uniform(-np.pi, np.pi)), rng.uniform(-5, 5, 3) * [1, 1, 0.05])
        clouds.append(synth.transform_points(T, scans[i % len(scans)]))
    pts = np.concatenate(clouds)
    del clouds
    off = np.arange(len(owned) + 1, dtype=np.int64) * N_SCAN
    g.enableTiming(True)
    out = {}
    for rep in range(3):
        g.timings(reset=True)
        t0 = time.perf_counter()
        g.batchCovariances(pts, off)
        wall = time.perf_counter() - t0
        t = g.timings(reset=True)
        out = {"points": int(len(pts)), "keyframes": len(owned), "index_ms": t["index_ms"], "knn_ms": t["knn_ms"], "covariance_ms": t["covariance_ms"],
               "wall_ms_with_h2d_d2h": 1e3 * wall}
    g.enableTiming(False)
    n = len(pts)
    dev_ms = out["index_ms"] + out["knn_ms"] + out["covariance_ms"]
    out["covariance_mpts_s_device"] = n / (dev_ms * 1e-3) / 1e6
    k3 = BYTES["K3_cov_per_pt"] * n / (out["covariance_ms"] * 1e-3) / 1e9
    out["roofline_K3"] = {"bound": "hbm", "achieved": k3, "peak": hbm["gbs"], "unit": "GB/s", "frac": k3 / hbm["gbs"],
                          "traffic": ncu_traffic("bulk_K3"), "traffic_note": "ncu --set full capture of the 256-keyframe (16,777,216-point) launch, per point x this launch's points",
                          "algorithmic_bytes_per_pt": BYTES["K3_cov_per_pt"], "peak_source": hbm["source"]}
    if out["roofline_K3"]["traffic"] is not None:
        out["roofline_K3"]["traffic"] *= n / float(ncu_traffic("bulk_K3_points") or CFG3_KEYFRAMES * N_SCAN)
    return out


def bulk_covariance(g, scans, hbm, n_keyframes=64):
    """Bulk covariance build: n_keyframes x 65,536 points in one batched pass, every keyframe its own index
    (BASELINE config 3 shape). Reports K2+K3 throughput and K3's HBM roofline fraction."""
    from ngicp import synth
    rng = np.random.default_rng(99)
    clouds = []
    for i in range(n_keyframes):
        T = synth.se3((0, 0, rng.uniform(-np.pi, np.pi)), rng.uniform(-5, 5, 3) * [1, 1, 0.05])
        clouds.append(synth.transform_points(T, scans[i % len(scans)]))
    pts = np.concatenate(clouds)
    off = np.arange(n_keyframes + 1, dtype=np.int64) * N_SCAN
    g.enableTiming(True)
    out = {}
    for rep in range(3):
        g.timings(reset=True)
        t0 = time.perf_counter()
        g.batchCovariances(pts, off)
        wall = time.perf_counter() - t0
        t = g.timings(reset=True)
        out = {"points": int(len(pts)), "keyframes": n_keyframes, "index_ms": t["index_ms"], "knn_ms": t["knn_ms"], "covariance_ms": t["covariance_ms"],
               "wall_ms_with_h2d_d2h": 1e3 * wall}
    g.enableTiming(False)
    n = len(pts)
    dev_ms = out["index_ms"] + out["knn_ms"] + out["covariance_ms"]
    out["covariance_mpts_s_device"] = n / (dev_ms * 1e-3) / 1e6
    k3 = BYTES["K3_cov_per_pt"] * n / (out["covariance_ms"] * 1e-3) / 1e9
    tr = ncu_traffic("bulk_K3")
    out["roofline_K3"] = {"bound": "hbm", "achieved": k3, "peak": hbm["gbs"], "unit": "GB/s", "frac": k3 / hbm["gbs"],
                          "traffic": None if tr is None else tr * n / float(ncu_traffic("bulk_K3_points") or n),
                          "algorithmic_bytes_per_pt": BYTES["K3_cov_per_pt"], "peak_source": hbm["source"]}
    return out


def multi_sequence_probe(g, tgt, m4, d_scans, device, sequences=4, steps=24):
    """Capacity of ONE GPU when several independent sequences (robots / replays) register against the same resident
    submap at once: one handle + stream + host thread per sequence, the target tree shared (DLIO hands trees between
    NanoGICP objects the same way, odom.cc:992-998). A single sequence is latency-bound (the LM loop is serial), so the
    aggregate is what the hardware can actually sustain. Reported beside the headline, never instead of it."""
    import ngicp
    handles = [g]
    for _ in range(sequences - 1):
        h = configure(ngicp.NanoGICP(device))
        h.registerInputTarget(tgt)
        h.setTargetTree(g.target_kdtree_)
        h.setTargetCovariances(m4)
        handles.append(h)
    for h in handles:                                   # warm every handle (allocations, first-use costs)
        for i in range(2):
            ds = d_scans[i % len(d_scans)]
            h.setInputSourceDevice(ds.data_ptr(), ds.shape[0], token=ds); h.calculateSourceCovariances(); h.align()
    start = threading.Barrier(sequences + 1)

    def worker(h, k):
        start.wait()
        for i in range(steps):
            ds = d_scans[(i + k) % len(d_scans)]
            h.setInputSourceDevice(ds.data_ptr(), ds.shape[0], token=ds); h.calculateSourceCovariances(); h.align()
        h.synchronize()

    ths = [threading.Thread(target=worker, args=(h, k)) for k, h in enumerate(handles)]
    for t in ths:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return {"multi_sequence": {"sequences": sequences, "scans_per_s": sequences * steps / dt, "note": "wall clock, no L2 flush, device-resident scans; "
                               "independent sequences on one GPU sharing the resident submap"}}


def prefilter_probe(g, pinned, h_scans, with_cpu):
    """CropBox -> VoxelGrid -> setInputSource on the device for one raw 65,536-point scan (host buffer in,
    index out), beside the same two PCL filters restated on one CPU core (PCL's filters are single-threaded)."""
    crop = ([-1.0] * 3, [1.0] * 3, True)        # cfg/params.yaml: crop size 1.0
    leaf = (0.25, 0.25, 0.25)
    ts = []
    n_out = 0
    for i in range(8):
        hs = pinned[i % len(pinned)].numpy()
        hs[:] = h_scans[i % len(h_scans)]
        t0 = time.perf_counter()
        n_out = len(g.setInputSourceFiltered(hs, crop=crop, leaf=leaf))
        ts.append(time.perf_counter() - t0)
    out = {"prefilter_points_in": N_SCAN, "prefilter_points_out": int(n_out), "prefilter_ms_crop_voxel_index": 1e3 * float(np.median(ts[2:]))}
    if with_cpu:
        import oracle
        oc = oracle.CropBox(); oc.setNegative(True); oc.setMin(crop[0]); oc.setMax(crop[1])
        ov = oracle.VoxelGrid(); ov.setLeafSize(*leaf)
        t0 = time.perf_counter()
        for i in range(5):
            oc.setInputCloud(h_scans[i % len(h_scans)]); ov.setInputCloud(oc.filter()); ov.filter()
        out["prefilter_cpu_ms_crop_voxel"] = 1e3 * (time.perf_counter() - t0) / 5
    return out


def bulk_linearize(g, scans, hbm, n_scans=64):
    """64 scans registered against the resident 1M-point submap in ONE batched linearize (two launches): the streaming
    fused-linearisation kernel K4b at a size where HBM, not launch latency, is the bound."""
    from ngicp import synth
    rng = np.random.default_rng(7)
    clouds, Ts = [], []
    for i in range(n_scans):
        clouds.append(synth.transform_points(synth.random_se3(rng, 0.05, 0.5), scans[i % len(scans)]))
        Ts.append(synth.random_se3(rng, 0.02, 0.2))
    pts = np.concatenate(clouds)
    off = np.arange(n_scans + 1, dtype=np.int64) * N_SCAN
    g.setInputSourceBatch(pts, off)
    g.calculateSourceCovariances()
    g.enableTiming(True)
    out = {}
    for rep in range(3):
        g.timings(reset=True)
        e, H, b, nc = g.batchLinearize(np.stack(Ts))
        t = g.timings(reset=True)
        out = {"batch_scans": n_scans, "batch_points": int(len(pts)), "batch_correspond_ms": t["correspond_ms"], "batch_linearize_ms": t["linearize_ms"],
               "batch_matched_fraction": float(nc.sum() / len(pts))}
    g.enableTiming(False)
    n = len(pts)
    k4b = BYTES["K4b_lin_per_src_pt"] * n / (out["batch_linearize_ms"] * 1e-3) / 1e9
    out["roofline_K4b"] = {"bound": "hbm", "achieved": k4b, "peak": hbm["gbs"], "unit": "GB/s", "frac": k4b / hbm["gbs"], "traffic": ncu_traffic("bulk_K4b"),
                           "algorithmic_bytes_per_pt": BYTES["K4b_lin_per_src_pt"], "peak_source": hbm["source"],
                           "note": "unmatched points move 20 B instead of 84 B; bytes are counted as if every point matched"}
    out["batch_scans_per_s_linearize_only"] = n_scans / ((out["batch_correspond_ms"] + out["batch_linearize_ms"]) * 1e-3)
    return out


def peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return {"gbs": float(json.loads(p.read_text())["hbm_gbs"]), "source": "MEASURED_PEAKS.json (of measured)"}
        except Exception:
            pass
    return {"gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    from ngicp import sharding
    rank, local_rank, world = sharding.world()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "b200":
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", str(Path(__file__).resolve()),
                                   "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)])
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
