#!/usr/bin/env python
"""bench.py -- control-steps/s of the batched safety-filter solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5|loop2|loop4]

One "step" = one pass of the hot path over one batch of synthetic agents.  Default workload:
  N = 1   cfg2 = BASELINE.json configs[1] (1024 DynamicUnicycle2D agents x 16 circular obstacles, cbf_qp); the line also
          embeds `sub_records` for cfg3, cfg4 and cfg5 (each: value, kernel_ms, roofline, e2e) measured in the same run.
  N > 1   cfg5 = BASELINE.json configs[4]: 65536 mixed du/kb/quad3d agents, mpc_cbf H=10, M=64 -- STRONG scaling: rank 0
          holds the batch, NCCL scatters it per model group, every rank solves its blocks, NCCL gathers U/status/iters/active
          (safe_control_b200.mixed.ShardedMixedMPCCBF), all inside the timed region.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.

  value     whole-job control-steps/s, inputs resident in HBM, CUDA events on the launch stream,
            max over ranks.  A pool of distinct batches larger than 2x L2 is cycled so no step
            re-reads L2-resident inputs.
  e2e       same metric through the host-pointer C-ABI call (scb_*_solve_host): pinned host
            inputs -> H2D -> kernel -> D2H of (U, status, active) every step.
  roofline  algorithmic bytes per launch / CUDA-event kernel time vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle port (oracle/, numpy, one agent at a time like the reference's
            control_step loop) timed on a bounded sample on this box's host cores.
  --impl reference   the reference-equivalent CPU path (same oracle port, all host cores).
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

WORKLOADS = {
    # name: (model, controller, N agents, M obstacles, horizon, dynamic obstacles)
    "cfg2": dict(model="DynamicUnicycle2D", controller="cbf_qp", N=1024, M=16, H=0, dynamic=False,
                 desc="1024 DynamicUnicycle2D agents, cbf_qp, 16 circular obstacles"),
    "cfg3": dict(model="DynamicUnicycle2D", controller="mpc_cbf", N=4096, M=16, H=8, dynamic=False,
                 desc="4096 DynamicUnicycle2D agents, mpc_cbf horizon 8, 16 obstacles"),
    "cfg4": dict(model="KinematicBicycle2D_C3BF", controller="optimal_decay_cbf_qp", N=8192, M=32, H=0, dynamic=True,
                 desc="8192 KinematicBicycle2D_C3BF agents, optimal_decay_cbf_qp, 32 dynamic obstacles"),
    # config 5: 65536 mixed-model agents (1/3 DU, 1/3 KB, 1/3 Quad3D), the WHOLE batch sharded over the ranks (strong scaling)
    "cfg5": dict(model="mixed", controller="mpc_cbf", N=65536, M=64, H=10, dynamic=False,
                 desc="65536 mixed-model agents (du/kb/quad3d), mpc_cbf N=10, 64 obstacles, sharded across the ranks"),
}
WORKLOADS["loop2"] = dict(model="DynamicUnicycle2D", controller="cbf_qp", N=1024, M=16, H=0, dynamic=False, loop=True,
                          desc="closed loop of config 2 (SURVEY 8f-1): 1024 DynamicUnicycle2D agents x 16 obstacles, full "
                               "control_step() per agent on the device (state machine, selection, nominal input, cbf_qp, collision, step)")
WORKLOADS["loop4"] = dict(model="KinematicBicycle2D_C3BF", controller="optimal_decay_cbf_qp", N=8192, M=32, H=0, dynamic=True,
                          loop=True, desc="closed loop of config 4: 8192 KinematicBicycle2D_C3BF agents, optimal_decay_cbf_qp, "
                                          "32 moving obstacles, full control_step() on the device")
WORKLOADS["backup"] = dict(model="DoubleIntegrator2D", controller="backup_cbf_qp", N=65536, M=1, H=120, dynamic=True,
                           desc="SURVEY 8f-3: 65536 DoubleIntegrator2D agents in the evade scene, Backup-CBF QP "
                                "(120 backup steps with forward-difference sensitivities -> 120 rows + 4 box rows x 2 inputs), 1 moving obstacle each")
WORKLOADS["gatekeeper"] = dict(model="DoubleIntegrator2D", controller="gatekeeper", N=65536, M=1, H=120, dynamic=True,
                               desc="SURVEY 8f-4: 65536 DoubleIntegrator2D agents in the evade scene, gatekeeper shield: per step 22 candidates "
                                    "(nominal prefix of 100, 95, .. 0 steps + 120-step backup rollout) validated against walls + bullet")
WORKLOADS["mps"] = dict(WORKLOADS["gatekeeper"], controller="mps",
                        desc="SURVEY 8f-4: 65536 DoubleIntegrator2D agents in the evade scene, MPS shield: one candidate (1 nominal step + 120-step backup rollout) per step")
L2_BYTES = 126e6
MIXED = ("DynamicUnicycle2D", "KinematicBicycle2D", "Quad3D")


def algorithmic_bytes(w):
    """bytes per agent-step on the contract layout (f64, per-agent lists)."""
    nx, nu, M = (12, 4, w["M"]) if w["model"] == "Quad3D" else ((2, 2, w["M"]) if w["model"] == "SingleIntegrator2D" else (4, 2, w["M"]))
    if w["controller"] == "mpc_cbf":
        ng = 3 if w["model"] == "Quad3D" else 2
        return 8 * (nx + ng + nu + 7 * M + nu) + 12
    return 8 * (nx + nu + 7 * M + nu) + 4 + 8 * ((M + 2 * nu + 63) // 64)


def _profile_json(names):
    for n in names:
        try:
            with open(os.path.join(ROOT, "profiles", n)) as f:
                return json.load(f)
        except Exception:
            continue
    return {}


def ncu_traffic(workload):
    """dram__bytes_read+write per launch from the committed ncu capture of this kernel (profiles/r2_traffic.json)."""
    return _profile_json(["r2_traffic.json", "r1_traffic.json"]).get(workload)


def counted_flops(key):
    """FP64 flops per agent-solve (2 x DFMA + DADD + DMUL thread instructions / agents) counted by ncu on this workload's
    kernel: profiles/r2_flops.json, written by tools/count_flops.py from the capture named there."""
    return _profile_json(["r2_flops.json"]).get(key)


_FP64_PEAK = {}


def fp64_peak():
    """measured DFMA throughput of this device, TFLOP/s (scb_measure_fp64_peak: 16 independent chains per thread)"""
    import ctypes as C
    import torch
    from safe_control_b200._lib import lib, check
    dev = torch.cuda.current_device()
    if dev not in _FP64_PEAK:
        v = C.c_double()
        check(lib().scb_measure_fp64_peak(C.byref(v), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "scb_measure_fp64_peak")
        _FP64_PEAK[dev] = float(v.value)
    return _FP64_PEAK[dev], "measured (scb_measure_fp64_peak: DFMA microbenchmark in libscb.so, this run)"


def latency_floor():
    """us per launch of an empty / one-round-trip / two-round-trip kernel of cfg2's geometry inside a CUDA graph"""
    import ctypes as C
    from safe_control_b200._lib import lib, check
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    check(lib().scb_measure_latency_floor(C.byref(a), C.byref(b), C.byref(c), None), "scb_measure_latency_floor")
    return {"empty_kernel_us": a.value, "one_dependent_dram_trip_us": b.value, "two_dependent_dram_trips_us": c.value,
            "how": "scb_measure_latency_floor: 200 launches of 256 CTAs x 128 threads in one CUDA graph, best of 5 replays; "
                   "the chase kernels read a 512 MB table (4x L2) at per-launch offsets"}


def fp64_roofline(flops_per_step, k_ms, extra=None):
    """roofline object of an MPC workload: bound by FP64 issue / latency, not by HBM"""
    peak, src = fp64_peak()
    ach = (flops_per_step / (k_ms * 1e-3) / 1e12) if flops_per_step else None
    r = {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
         "traffic": None, "peak_source": src, "kernel_ms": k_ms,
         "flops_source": "ncu-counted 2*DFMA + DADD + DMUL thread instructions per agent (profiles/r2_flops.json) x agents per step"
                         if flops_per_step else "profiles/r2_flops.json missing: flops not counted",
         "note": "iterative interior-point NLP solve per agent: FP64 latency / issue bound, HBM traffic negligible (DESIGN.md 3.3)"}
    if extra:
        r.update(extra)
    return r


def timed_replays(graph, steps, torch, min_ms=50.0):
    """Replay a captured K-step graph until at least `min_ms` of device time is inside ONE event pair
    -> (ms per step, replays, total ms).  A 20-step x 5 us region is too thin a basis for a headline."""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); graph.replay(); b.record(); torch.cuda.synchronize()
    est = max(a.elapsed_time(b), 1e-3)
    reps = max(1, int(np.ceil(min_ms / est)))
    a.record()
    for _ in range(reps):
        graph.replay()
    b.record(); torch.cuda.synchronize()
    tot = a.elapsed_time(b)
    return tot / (reps * steps), reps, tot


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------- CPU arm
def _cpu_solve_range(args):
    """Worker: oracle port over a slice of agents, one at a time (the reference's control flow)."""
    w, sc, lo, hi = args[:4]
    budget = args[4] if len(args) > 4 else None          # seconds this worker may spend (MPC: SLSQP can stall for minutes)
    if budget != "inproc":
        try:                                             # pool worker: one core each -- BLAS / OpenMP pools of 16 processes x 16
            from threadpoolctl import threadpool_limits  # threads would fight over the same cores (observed: 60x slower)
            threadpool_limits(1)
        except Exception:
            pass
    from oracle.controllers import OracleCBFQP, OracleOptimalDecayCBFQP
    M = w["M"]
    t0 = time.perf_counter()
    if w["controller"] == "cbf_qp":
        ctrl = OracleCBFQP(sc["spec"], num_obs=M)
        for i in range(lo, hi):
            k = int(sc["nobs"][i])
            ctrl.solve(sc["X"][i], sc["U_ref"][i], sc["OBS"][i][:k])
    elif w["controller"] == "optimal_decay_cbf_qp":
        ctrl = OracleOptimalDecayCBFQP(sc["spec"])
        for i in range(lo, hi):
            k = int(sc["nobs"][i])
            ctrl.solve(sc["X"][i], sc["U_ref"][i], sc["OBS"][i][0] if k else None)   # lists are distance-sorted
    else:
        import torch
        torch.set_num_threads(1)               # one core per worker process (the pool already uses every host core)
        from oracle.mpc_cbf import OracleMPCCBF
        ctrl = OracleMPCCBF(sc["spec"], num_obs=M, horizon=w["H"])
        done = 0
        for i in range(lo, hi):
            k = int(sc["nobs"][i])
            # (iteration cap 40 instead of the tests' 400 -- converging runs take 15-30 -- because a stalled SLSQP run costs
            #  minutes at config-5 size; IPOPT in the reference likewise returns its last iterate at its own limit)
            ctrl.solve(sc["X"][i], sc["goal"][i], sc["u_prev"][i], sc["OBS"][i][:k], maxiter=40)
            done += 1
            if isinstance(budget, (int, float)) and time.perf_counter() - t0 > budget:
                break
        return done, time.perf_counter() - t0
    return hi - lo, time.perf_counter() - t0


_POOL = {}


def _close_pools():
    for pool in _POOL.values():
        pool.close(); pool.join()
    _POOL.clear()


def cpu_rate(w, sc, n_agents, procs, offset=0, budget=None):
    """agent-steps/s of the oracle port on `procs` host processes over agents [offset, offset + n_agents); `budget` =
    seconds after which a worker stops taking new agents (it still counts what it finished)."""
    n_tot = sc["X"].shape[0]
    offset = offset % max(n_tot - n_agents + 1, 1)
    n_agents = min(n_agents, n_tot)
    if procs <= 1:
        n, dt = _cpu_solve_range((w, sc, offset, offset + n_agents, "inproc"))
        return n / dt, n, dt
    import multiprocessing as mp
    if procs not in _POOL:
        _POOL[procs] = mp.get_context("fork").Pool(procs)      # created once; workers import the oracle lazily
    bounds = offset + np.linspace(0, n_agents, procs + 1).astype(int)
    small = {k: sc[k] for k in ("spec", "X", "U_ref", "OBS", "nobs", "goal", "u_prev")}
    t0 = time.perf_counter()
    if w["controller"] == "mpc_cbf":
        # solve times vary 10x between agents (SLSQP stalls on some): hand agents out one at a time so that no worker idles
        # behind a straggler; only the rows a task needs travel to the worker
        tasks = []
        for i in range(offset, offset + n_agents):
            one = {k: (small[k] if k == "spec" else small[k][i:i + 1]) for k in small}
            tasks.append((w, one, 0, 1, None))
        res = list(_POOL[procs].imap_unordered(_cpu_solve_range, tasks, chunksize=1))
        dt = time.perf_counter() - t0
        return sum(r[0] for r in res) / dt, sum(r[0] for r in res), dt
    res = _POOL[procs].map(_cpu_solve_range, [(w, small, int(bounds[j]), int(bounds[j + 1]), budget) for j in range(procs)])
    dt = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    return n / dt, n, dt


def cpu_sample_size(w):
    return {"cbf_qp": 4096, "optimal_decay_cbf_qp": 8192, "mpc_cbf": 48}[w["controller"]]


def _oracle_loop_range(args):
    """Worker: the reference's control flow (oracle/tracking.py), one agent at a time, `steps` control steps each."""
    w, X0, scene, wps, spec, lo, hi, steps = args
    from oracle.tracking import OracleTrackingController
    t0 = time.perf_counter()
    n = 0
    for i in range(lo, hi):
        tc = OracleTrackingController(X0[i], spec, w["controller"], obs=scene.copy(), dynamic_obs=w["dynamic"])
        tc.set_waypoints(wps[i])
        for _ in range(steps):
            n += 1
            if tc.control_step() in (-1, -2):
                break
    return n, time.perf_counter() - t0


def loop_case(w, N, seed):
    """Seeded closed-loop scene (SURVEY 8d arena: side 4 sqrt(M), radii U[0.2, 0.6]) + 3 waypoints per agent."""
    rng = np.random.default_rng(seed)
    M = w["M"]
    L = 4.0 * np.sqrt(M)
    scene = np.zeros((M, 7))
    scene[:, 0:2] = rng.uniform(0, L, (M, 2)); scene[:, 2] = rng.uniform(0.2, 0.6, M)
    if w["dynamic"]:
        scene[:, 3:5] = rng.uniform(-0.5, 0.5, (M, 2))
    pos = np.empty((N, 2)); todo = np.arange(N)
    while todo.size:
        cand = rng.uniform(0, L, (todo.size, 2))
        ok = ((np.sqrt(((cand[:, None] - scene[None, :, :2]) ** 2).sum(-1)) - scene[None, :, 2] - 0.8) > 0).all(1)
        pos[todo[ok]] = cand[ok]; todo = todo[~ok]
    v = rng.uniform(0.2, 1.0, N)
    X0 = np.hstack([pos, rng.uniform(-np.pi, np.pi, (N, 1)), v[:, None]])
    wps = np.zeros((N, 4, 3))
    wps[:, 0, :2] = pos
    wps[:, 1:, :2] = np.clip(pos[:, None, :] + np.cumsum(rng.uniform(-0.35 * L, 0.35 * L, (N, 3, 2)), axis=1), 0, L)
    return X0, scene, wps


def reference_loop(args, w, procs):
    """--impl reference for the closed-loop workloads: oracle/tracking.py, one agent at a time per process."""
    import multiprocessing as mp
    per_proc, n_s = 2, 25
    n_agents = per_proc * procs
    X0, scene, wps = loop_case(w, n_agents * (args.steps + 3), 1234)
    spec = {"model": w["model"], "num_constraints": w["M"]}
    pool = mp.get_context("fork").Pool(procs)
    def one(k):
        lo = k * n_agents
        jobs = [(w, X0, scene, wps, spec, lo + j * per_proc, lo + (j + 1) * per_proc, n_s) for j in range(procs)]
        t0 = time.perf_counter()
        res = pool.map(_oracle_loop_range, jobs)
        return sum(r[0] for r in res), time.perf_counter() - t0
    for k in range(max(1, min(args.warmup, 2))):
        one(k)
    n_tot, t_tot, done = 0, 0.0, 0
    for k in range(args.steps):
        n, dt = one(k + 2)
        n_tot += n; t_tot += dt; done += 1
        if t_tot > 150.0:
            break
    pool.close(); pool.join()
    val = n_tot / t_tot
    sample = f"{done} steps x {n_agents} agents x up to {n_s} control steps through oracle/tracking.py, {procs} processes"
    print(json.dumps({
        "impl": "reference", "metric": "control-steps/sec (batched QP solves/s)", "value": val, "unit": "control-steps/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(done, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {w['desc']}", "obstacles": w["M"]},
        "cpu_baseline": {"value": val, "unit": "control-steps/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "control-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference's cvxpy/GUROBI stack is not installable here; this is the oracle restatement of tracking.py:control_step",
    }))


def run_loop(args, w, rank, world, local_rank):
    """Closed loop on the device: one step = scb_control_step over N agents (3-4 launches)."""
    import torch
    import torch.distributed as dist
    from safe_control_b200 import BatchedTrackingController
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    N, M = w["N"], w["M"]
    X0, scene, wps = loop_case(w, N, 1234 + rank)
    spec = {"model": w["model"], "num_constraints": M}
    mk = lambda: BatchedTrackingController(X0, spec, {"pos": w["controller"]}, obs=scene, dynamic_obs=w["dynamic"], device=dev)
    tc = mk(); tc.set_waypoints(wps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tc.run_steps(args.warmup)
    barrier()
    graph = torch.cuda.CUDAGraph()
    l0 = tc.launches
    with torch.cuda.graph(graph):
        tc.run_steps(args.steps)
    launches = tc.launches - l0
    # every timed replay starts from the same tracker state (after the warm-up steps): snapshot / restore
    keys = ("X", "yaw", "sm", "wp_idx", "goal", "has_goal", "u_att", "u_prev", "ret", "done", "nsteps", "SCENE")
    snap = {k: tc.buffers()[k].clone() for k in keys}
    def restore():
        for k in keys:
            tc.buffers()[k].copy_(snap[k])
    graph.replay(); restore(); barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    reps = []
    for _ in range(5):
        restore(); barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b.record(); barrier()
        reps.append(a.elapsed_time(b))
    ms = float(np.median(reps))
    clocks = sampler.stop() if rank == 0 else None
    bufs = tc.buffers()
    active_frac = float((bufs["nsteps"] - snap["nsteps"]).float().mean()) / args.steps
    rets = bufs["ret"].cpu().numpy(); done = bufs["done"].cpu().numpy()
    # eager: one C call launching 3-4 kernels per step, no graph
    restore(); barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(); tc.run_steps(args.steps); g1.record(); torch.cuda.synchronize()
    eager_ms = g0.elapsed_time(g1)
    # the same K steps through n x scb_control_step (3-4 launches per step), for comparison with the fused launch
    os.environ["SCB_TRACK_FUSED"] = "0"
    restore(); barrier()
    g0.record(); tc.run_steps(args.steps); g1.record(); torch.cuda.synchronize()
    per_step = {"value": world * N * args.steps / (g0.elapsed_time(g1) * 1e-3), "unit": "control-steps/s",
                "how": "SCB_TRACK_FUSED=0: pre / solve / post kernels per step, eager launches (this rank)"}
    del os.environ["SCB_TRACK_FUSED"]
    # end to end: initial state from pinned host memory -> device, run K steps, final state + return codes back
    host = {k: snap[k].cpu().pin_memory() for k in keys}
    out_h = {k: torch.empty_like(host[k]).pin_memory() for k in ("X", "ret", "nsteps")}
    def e2e_run():
        for k in keys:
            tc.buffers()[k].copy_(host[k], non_blocking=True)
        tc.run_steps(args.steps)
        for k in out_h:
            out_h[k].copy_(tc.buffers()[k], non_blocking=True)
        torch.cuda.synchronize()
    e2e_run(); barrier()
    t0 = time.perf_counter(); e2e_run(); e2e_ms = (time.perf_counter() - t0) * 1e3
    h2d = sum(host[k].numel() * host[k].element_size() for k in keys)
    d2h = sum(v.numel() * v.element_size() for v in out_h.values())
    tm = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_src = hbm_peak()
        nx, nu = 4, 2
        # tracker state read + written once per agent-step; the selected obstacle rows are written by the pre kernel and
        # read by the solve kernel (they stay in L2); the shared scene is read from L2
        B = 2 * 8 * (nx + 1 + 2 + 1 + nu + nu) + 2 * 4 * 6 + 2 * 56 * M + 24
        k_ms = float(tm[0]) / args.steps
        out = {
            "metric": "control-steps/sec (batched QP solves/s)", "value": world * N * args.steps / (float(tm[0]) * 1e-3),
            "unit": "control-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": k_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{w['name']}: {w['desc']}", "agents_per_step_per_gpu": N, "obstacles": M, "scene_obstacles": M,
                       "waypoints_per_agent": 3, "launch": f"K steps = {launches} launch(es) (QP controllers: one fused kernel keeps every agent in a warp for all K steps) captured in a CUDA graph; every replay restarts from the same tracker state",
                       "l2_policy": "closed loop: the state produced by step k is the input of step k+1 (nothing is re-read from a warm copy); agents_active_frac = share of agent-steps not yet frozen by a -1/-2 return",
                       "agents_active_frac": active_frac, "final_ret": {"0": int((rets == 0).sum()), "-1": int((rets == -1).sum()), "-2": int((rets == -2).sum())},
                       "parallelism": f"agents sharded, {world} rank(s), no data-path collective"},
            "e2e": {"value": world * N * args.steps / (float(tm[1]) * 1e-3), "unit": "control-steps/s", "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, "steps_timed": args.steps,
                    "how": "whole run_all_steps: tracker state pinned host -> device, K control steps on the device, final X / ret / nsteps -> pinned host, sync"},
            "gpu_launches": launches,
            "eager": {"value": world * N * args.steps / (eager_ms * 1e-3), "unit": "control-steps/s", "how": "scb_run_all_steps without the CUDA graph"},
            "per_step_path": per_step,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": B * N / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": B * N / (k_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": k_ms,
                         "algorithmic_bytes_per_agent": B, "agents_per_launch": N,
                         "note": "closed loop: step k+1 needs step k, so one agent-warp runs its K steps back to back (latency-bound by construction); per-step traffic stays in L1/L2"},
        }
        if not args.no_cpu:
            n_a, n_s = 24, 50
            n, dt = _oracle_loop_range((w, X0, scene, wps, spec, 0, n_a, n_s))
            out["cpu_baseline"] = {"value": n / dt, "unit": "control-steps/s", "cores": 1, "kind": "port", "host_cores_available": os.cpu_count(),
                                   "sample": f"{n_a} agents x up to {n_s} control steps of the same scene through oracle/tracking.py (one agent at a time), {dt:.1f} s"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------- device arms
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_single(args, w, cx, steps, warmup, sub=False):
    """One (model, controller) workload, agents resident per rank (weak scaling when world > 1).  -> record (rank 0)."""
    torch = cx.torch
    from safe_control_b200 import BatchedCBFQP, BatchedOptimalDecayCBFQP, BatchedMPCCBF, HostContext, scenes
    rank, world, dev = cx.rank, cx.world, cx.dev
    N, M = w["N"], w["M"]
    B = algorithmic_bytes(w)
    od = w["controller"] == "optimal_decay_cbf_qp"
    mpc = w["controller"] == "mpc_cbf"

    # ---- input pool: P distinct batches, > 2x L2 in total (4 for the compute-bound MPC) ----
    P = int(np.ceil(2.2 * L2_BYTES / (B * N))) if not mpc else 4
    sc = scenes.make_scene(w["model"], N * P, M, seed=1234 + rank, dynamic=w["dynamic"], optimal_decay=od)
    spec = sc["spec"]
    t = lambda a: torch.from_numpy(a).to(dev)
    Xp, Urp, OBSp, nobsp = t(sc["X"]).view(P, N, -1), t(sc["U_ref"]).view(P, N, -1), t(sc["OBS"]).view(P, N, M, 7), t(sc["nobs"]).view(P, N)
    goalp, uprevp = t(sc["goal"]).view(P, N, -1), t(sc["u_prev"]).view(P, N, -1)

    if w["controller"] == "cbf_qp":
        ctrl = BatchedCBFQP(spec, num_obs=M)
        outs = (torch.empty((N, 2), dtype=torch.float64, device=dev), torch.empty(N, dtype=torch.int32, device=dev),
                torch.empty((N, ctrl.words), dtype=torch.int64, device=dev))
        step = lambda k: ctrl.solve(Xp[k % P], Urp[k % P], OBSp[k % P], nobsp[k % P], out=outs)
    elif od:
        ctrl = BatchedOptimalDecayCBFQP(spec, num_obs=M)
        step = lambda k: ctrl.solve(Xp[k % P], Urp[k % P], OBSp[k % P], nobsp[k % P])
    else:
        ctrl = BatchedMPCCBF(spec, num_obs=M, horizon=w["H"])
        step = lambda k: ctrl.solve(Xp[k % P], goalp[k % P], uprevp[k % P], OBSp[k % P], nobsp[k % P], want_active=True)

    # ---- device-resident throughput ("value"): K steps captured once into a CUDA graph, the replay timed with CUDA
    # events on the launching stream and repeated until >= 50 ms are inside the event pair ----
    for k in range(warmup):
        step(k)
    cx.barrier()
    graph = torch.cuda.CUDAGraph()
    l0 = ctrl.launches
    with torch.cuda.graph(graph):
        for k in range(steps):
            last = step(warmup + k)
    launches = ctrl.launches - l0
    graph.replay()                       # warm replay (instantiation, first-touch)
    cx.barrier()
    sampler = ClockSampler(cx.local_rank)
    if rank == 0 and not sub:
        sampler.start()
    cx.barrier()
    ms_step, reps, tot_ms = timed_replays(graph, steps, torch, 50.0 if not mpc else 0.0)
    cx.barrier()
    k_list = [timed_replays(graph, steps, torch, 20.0 if not mpc else 0.0)[0] for _ in range(3 if mpc else 5)]
    k_ms = float(np.median(k_list))
    clocks = sampler.stop() if (rank == 0 and not sub) else None
    solver = None
    if mpc:
        solver = {"optimal_frac": float((last["status"] == 0).float().mean()), "iters_mean": float(last["iters"].float().mean()),
                  "iters_max": int(last["iters"].max()),
                  "active_rows_mean": float(sum(((last["active"] >> b) & 1).sum() for b in range(64)).item()) / N}
    eager = None
    if not sub:                          # eager (no graph): Python + ctypes + launch per step
        cx.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for k in range(steps):
            step(warmup + k)
        g1.record(); torch.cuda.synchronize()
        eager = g0.elapsed_time(g1)

    # asymptote of the same kernel on a batch that fills the machine (not the headline; explains it)
    big = None
    if not mpc:
        NB = 1 << 20
        rp = NB // (N * P) + 1
        Xb = Xp.reshape(N * P, -1).repeat(rp, 1)[:NB].contiguous(); Ub = Urp.reshape(N * P, -1).repeat(rp, 1)[:NB].contiguous()
        Ob = OBSp.reshape(N * P, M, 7).repeat(rp, 1, 1)[:NB].contiguous(); nb = nobsp.reshape(-1).repeat(rp)[:NB].contiguous()
        for _ in range(3):
            ctrl.solve(Xb, Ub, Ob, nb)
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(20):
            ctrl.solve(Xb, Ub, Ob, nb)
        b1.record(); torch.cuda.synchronize()
        big_ms = b0.elapsed_time(b1) / 20
        peak_, _ = hbm_peak()
        big = dict(agents=NB, ms_per_launch=big_ms, control_steps_per_s=NB / big_ms * 1e3, gbs=B * NB / big_ms / 1e6,
                   frac=B * NB / big_ms / 1e6 / peak_, traffic=ncu_traffic(w["name"] + "_large_batch"),
                   how="20 launches back to back over a 1 Mi-agent batch (1 GB of inputs > L2), CUDA events")
        del Xb, Ub, Ob, nb

    # ---- activity mix of the timed inputs (iteration counts depend on it) ----
    mix = None
    if w["controller"] == "cbf_qp":
        st_all, cbf_act, box_act = [], [], []
        for k in range(min(P, 16)):
            U, st, act = ctrl.solve(Xp[k], Urp[k], OBSp[k], nobsp[k])
            st_all.append(st.clone()); a = act[:, 0].clone()
            cbf_act.append((a & ((1 << M) - 1)) != 0); box_act.append((a >> M) != 0)
        st_all = torch.cat(st_all); cbf_act = torch.cat(cbf_act); box_act = torch.cat(box_act)
        mix = dict(infeasible=float((st_all == 1).float().mean()), cbf_active=float(cbf_act.float().mean()),
                   box_active=float(box_act.float().mean()))

    # ---- end to end through the host-pointer C ABI ----
    ctx = HostContext(cx.local_rank)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    P_h = min(P, 32)
    hX = [pin(sc["X"][k * N:(k + 1) * N]) for k in range(P_h)]
    hU = [pin(sc["U_ref"][k * N:(k + 1) * N]) for k in range(P_h)]
    hO = [pin(sc["OBS"][k * N:(k + 1) * N]) for k in range(P_h)]
    hn = [pin(sc["nobs"][k * N:(k + 1) * N]) for k in range(P_h)]
    hg = [pin(sc["goal"][k * N:(k + 1) * N]) for k in range(P_h)]
    hp = [pin(sc["u_prev"][k * N:(k + 1) * N]) for k in range(P_h)]
    if w["controller"] == "cbf_qp":
        ho = (pin(np.empty((N, 2))), pin(np.empty(N, np.int32)), pin(np.empty((N, ctrl.words), np.uint64)))
        estep = lambda k: ctx.cbfqp_solve(ctrl.params, M, hX[k % P_h], hU[k % P_h], hO[k % P_h], hn[k % P_h], out=ho)
        h2d = N * (4 + 2) * 8 + N * M * 56 + N * 4; d2h = N * 2 * 8 + N * 4 + N * ctrl.words * 8
    elif od:
        ho = (pin(np.empty((N, 2))), pin(np.empty((N, 2))), pin(np.empty(N, np.int32)), pin(np.empty(N, np.int32)),
              pin(np.empty(N, np.uint64)))
        estep = lambda k: ctx.odcbf_solve(ctrl.params, M, hX[k % P_h], hU[k % P_h], hO[k % P_h], hn[k % P_h], out=ho)
        h2d = N * (4 + 2) * 8 + N * M * 56 + N * 4; d2h = N * 2 * 8 * 2 + N * 4 * 2 + N * 8
    else:
        estep = lambda k: ctx.mpccbf_solve(ctrl.params, M, w["H"], hX[k % P_h], hg[k % P_h], hp[k % P_h], hO[k % P_h], hn[k % P_h],
                                           want_active=True)
        h2d = N * (ctrl.nx + ctrl.ngoal + ctrl.nu) * 8 + N * M * 56 + N * 4
        d2h = N * ctrl.nu * 8 + N * 4 * 2 + N * 8 + N * ctrl.active_words * 8
    # (QP workloads: at least 200 steps whatever --steps says -- 20 steps x 56 us is a 1 ms sample -- and the median of three
    #  passes; the MPC call is milliseconds per step)
    e_steps = max(200, min(steps, 400)) if not mpc else max(10, min(steps, 20))

    def time_e2e():
        for k in range(min(warmup, 10)):
            estep(k)
        cx.barrier()
        t0 = time.perf_counter()
        for k in range(e_steps):
            estep(k)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    e2e_s = float(np.median([time_e2e() for _ in range(3 if not mpc else 1)]))    # the library's own choice (zero-copy for page-locked buffers <= 32 MB)
    e2e_launches = ctx.launches
    staged = None
    if not sub and not mpc:
        os.environ["SCB_HOST_PATH"] = "staged"   # same call, forced through explicit H2D / D2H staging copies
        staged = time_e2e()
        del os.environ["SCB_HOST_PATH"]
    ctx.close()

    ms_max, e2e_ms_max = cx.max_over_ranks([ms_step, e2e_s * 1e3 / e_steps])
    if rank != 0:
        return None
    if mpc:
        fl = counted_flops(w["name"])
        roof = fp64_roofline(fl * N if fl else None, k_ms, {"agents_per_launch": N, "flops_per_agent": fl,
                                                            "hbm_algorithmic_bytes_per_agent": B})
    else:
        peak, peak_src = hbm_peak()
        achieved = B * N / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(w["name"]), "peak_source": peak_src, "kernel_ms": k_ms,
                "kernel_ms_how": "CUDA-graph replay of the timed steps / steps (median of 5 regions of >= 20 ms each), includes the inter-kernel dependency gap",
                "algorithmic_bytes_per_agent": B, "agents_per_launch": N,
                "latency_floor": latency_floor() if (w["name"] == "cfg2" and not sub) else None,
                "note": f"one launch covers only {N} agents ({B * N / 1e6:.1f} MB): latency-bound; large_batch shows the same kernel on 1M agents",
                "large_batch": big}
    out = {
        "metric": "control-steps/sec (batched QP solves/s)",
        "value": world * N / (ms_max * 1e-3), "unit": "control-steps/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {w['desc']}", "agents_per_step_per_gpu": N, "obstacles": M,
                   "horizon": w["H"], "scene": "SURVEY 8d generator, seed 1234+rank, num_constraints=M",
                   "launch": f"K = {steps} steps captured in one CUDA graph, replayed {reps}x inside one CUDA-event pair ({tot_ms:.1f} ms timed)",
                   "steps_timed": reps * steps,
                   "l2_policy": f"inputs larger than L2: {P} distinct batches ({P * N * B / 1e6:.0f} MB) cycled" if not mpc
                                else f"compute-bound NLP solves; {P} distinct batches cycled",
                   "parallelism": f"agents sharded, {world} rank(s), no data-path collective", "activity_mix": mix, "solver": solver},
        "e2e": {"value": world * N / (e2e_ms_max * 1e-3), "unit": "control-steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps_timed": e_steps,
                "how": ("scb_*_solve_host per step on page-locked host arrays; the QP calls run zero-copy (kernel reads inputs / writes "
                        "U,status,active over PCIe through the mapped host buffers, then sync); the MPC call stages H2D -> kernel -> D2H")},
        "gpu_launches": launches * reps,
        "e2e_gpu_launches": e2e_launches,
        "roofline": roof,
    }
    if staged is not None:
        out["e2e"]["staged"] = {"value": world * N * e_steps / staged, "how": "SCB_HOST_PATH=staged: pinned host arrays -> cudaMemcpyAsync H2D -> kernel -> D2H -> sync (this rank)"}
    if eager is not None:
        out["eager"] = {"value": world * N * steps / (eager * 1e-3), "unit": "control-steps/s",
                        "how": "same steps without the CUDA graph: one Python->ctypes->launch per step (host-launch bound)"}
    if clocks is not None:
        out["clocks"] = clocks
    if not args.no_cpu and not sub:
        n_s = cpu_sample_size(w)
        rate, n, dt = cpu_rate(w, sc, n_s, 1)
        out["cpu_baseline"] = {"value": rate, "unit": "control-steps/s", "cores": 1, "kind": "port",
                               "host_cores_available": os.cpu_count(),
                               "sample": f"first {n} agents of the timed scene, oracle port (numpy, one agent at a time), {dt:.1f} s"}
    return out


def run_cfg5(args, w, cx, steps, warmup, sub=False):
    """BASELINE config 5, strong scaling: rank 0 holds the whole 65536-agent mixed batch (three model groups);
    per step  NCCL scatter -> one launch per group on concurrent streams -> NCCL gather of U / status / iters / active,
    all inside the timed region (safe_control_b200.mixed.ShardedMixedMPCCBF).  world == 1: same code, no collective."""
    torch = cx.torch
    from safe_control_b200 import scenes
    from safe_control_b200.mixed import ShardedMixedMPCCBF, MixedMPCCBF, split_counts
    rank, world, dev = cx.rank, cx.world, cx.dev
    N, M, H = w["N"], w["M"], w["H"]
    counts = split_counts(N, len(MIXED))
    specs = [scenes.default_spec(m) for m in MIXED]
    keys = ("X", "goal", "u_prev", "OBS", "nobs")
    P = 2
    scs = ins = None
    if rank == 0:
        scs = [[scenes.make_scene(m, c, M, seed=1234 + 101 * q + 17 * g) for g, (m, c) in enumerate(zip(MIXED, counts))] for q in range(P)]
        ins = [[{k: torch.from_numpy(sc[k]).to(dev) for k in keys} for sc in row] for row in scs]
    sh = ShardedMixedMPCCBF(specs, counts, M, H, dev, want_active=True)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(k, marks=None):
        blocks = sh.scatter(ins[k % P] if rank == 0 else None)
        if marks is not None:
            marks[0].record()
        outs = sh.solve_local(blocks)
        if marks is not None:
            marks[1].record()
        return sh.gather(outs), outs

    for k in range(max(warmup, 1)):
        step(k)
    cx.barrier()
    sampler = ClockSampler(cx.local_rank)
    if rank == 0 and not sub:
        sampler.start()
    l0 = sh.launches
    marks = [(ev(), ev(), ev(), ev()) for _ in range(steps)]
    cx.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for k in range(steps):
        marks[k][0].record()
        res, outs = step(k + warmup, marks[k][1:3])
        marks[k][3].record()
    e1.record()
    cx.barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = sh.launches - l0
    clocks = sampler.stop() if (rank == 0 and not sub) else None
    ph = np.array([[m[0].elapsed_time(m[1]), m[1].elapsed_time(m[2]), m[2].elapsed_time(m[3])] for m in marks]).mean(axis=0)
    stat = None
    if rank == 0:
        stat = {m: {"optimal_frac": float((o["status"] == 0).float().mean()), "iters_mean": float(o["iters"].float().mean()),
                    "iters_max": int(o["iters"].max())} for m, o in zip(MIXED, res)}

    # ---- end to end: pinned host inputs on rank 0 -> H2D -> scatter -> solve -> gather -> D2H of U / status / active ----
    e_steps = max(1, min(steps, 3))
    h2d = d2h = 0
    if rank == 0:
        hin = [{k: torch.from_numpy(sc[k]).pin_memory() for k in keys} for sc in scs[0]]
        din = [{k: torch.empty_like(v, device=dev) for k, v in g.items()} for g in hin]
        hout = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in g.items()} for g in res]
        h2d = sum(v.numel() * v.element_size() for g in hin for v in g.values())
        d2h = sum(v.numel() * v.element_size() for g in hout for v in g.values())

    def estep():
        if rank == 0:
            for g, d in zip(hin, din):
                for k in keys:
                    d[k].copy_(g[k], non_blocking=True)
        r = sh.solve(din if rank == 0 else None)
        if rank == 0:
            for g, o in zip(hout, r):
                for k in g:
                    g[k].copy_(o[k], non_blocking=True)
        torch.cuda.synchronize()

    estep(); cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        estep()
    cx.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e_steps

    # ---- the same batch on ONE GPU (rank 0 alone; the others wait): the base of the strong-scaling curve ----
    base = None
    if world > 1:
        if rank == 0:
            one = MixedMPCCBF(specs, M, H)
            one.solve(ins[0]); torch.cuda.synchronize()
            b0, b1 = ev(), ev()
            b0.record(); one.solve(ins[1]); b1.record(); torch.cuda.synchronize()
            base = {"n_gpus": 1, "ms_per_step": b0.elapsed_time(b1), "value": N / (b0.elapsed_time(b1) * 1e-3),
                    "how": "rank 0 alone solves the whole batch (no scatter / gather), one step after one warm-up, CUDA events"}
        cx.barrier()

    ms_max, e2e_max, sc_ms, so_ms, ga_ms = cx.max_over_ranks([ms, e2e_ms, ph[0], ph[1], ph[2]])
    if rank != 0:
        return None
    B = {"DynamicUnicycle2D": 3680, "KinematicBicycle2D": 3680, "Quad3D": 3784}
    fl = counted_flops("cfg5") or {}
    flops_step = sum(fl.get(m, 0) * c for m, c in zip(MIXED, counts)) if all(m in fl for m in MIXED) else None
    plan_bytes = (sum(pl.scatter_bytes() for pl in sh.plans), sum(pl.gather_bytes() for pl in sh.plans))
    out = {
        "metric": "control-steps/sec (batched QP solves/s)", "value": N / (ms_max * 1e-3),
        "unit": "control-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {w['desc']}", "agents_total": N, "agents_per_gpu": N / world,
                   "groups": dict(zip(MIXED, counts)), "obstacles": M, "horizon": H,
                   "scene": "SURVEY 8d generator per model group, seed 1234 (+101 per batch, +17 per group), num_constraints=M",
                   "launch": "per step: scatter (5 tensors x 3 groups, one coalesced NCCL point-to-point group), 3 x (key, counting sort, solve) launches on 3 streams, gather (4 tensors x 3 groups, one group); eager",
                   "l2_policy": f"compute-bound NLP solves; {P} distinct batches alternated ({plan_bytes[0] / 1e6:.0f} MB of inputs each)",
                   "parallelism": (f"strong scaling over {world} rank(s): NCCL scatter of each model group's rows from rank 0 -> per-rank solve -> "
                                   f"NCCL gather of U/status/iters/active to rank 0, inside the timed region" if world > 1 else
                                   "1 rank: whole batch on one GPU (scatter / gather degenerate to views)"),
                   "phases_ms": {"scatter_ms": sc_ms, "solve_ms": so_ms, "gather_ms": ga_ms, "how": "CUDA events per step on each rank, mean over steps, max over ranks"},
                   "phases_ms_rank0": {"scatter_ms": float(ph[0]), "solve_ms": float(ph[1]), "gather_ms": float(ph[2]),
                                       "how": "the same events on rank 0 (the source / sink of the data) alone: on the other ranks the scatter and gather "
                                              "phases also contain the wait for rank 0 and for the slowest rank, so the max over ranks above over-counts them"},
                   "scatter_bytes": plan_bytes[0], "gather_bytes": plan_bytes[1], "solver": stat},
        "e2e": {"value": N / (e2e_max * 1e-3), "unit": "control-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps_timed": e_steps,
                "how": "ShardedMixedMPCCBF.solve per step: pinned host batch on rank 0 -> H2D -> scatter -> solve -> gather -> D2H of U/status/iters/active -> sync; wall clock between barriers, max over ranks"},
        "gpu_launches": launches,
        "roofline": fp64_roofline(flops_step / world if flops_step else None, so_ms,
                                  {"kernel_ms_how": "solve phase (3 concurrent model-group launches) per step on the slowest rank, CUDA events",
                                   "flops_per_agent": fl or None, "hbm_algorithmic_bytes_per_agent": B, "per_gpu": True}),
    }
    if base is not None:
        out["strong_scaling_base"] = base
        out["strong_scaling_efficiency_vs_base"] = out["value"] / (world * base["value"])
        out["config"]["scaling_note"] = ("the N = 1 line of bench.py is config 2 (BASELINE configs[1]); the 1-GPU point of THIS workload is "
                                         "strong_scaling_base (same batch, rank 0 alone, measured in this run) and sub_records.cfg5 of the N = 1 line")
    if clocks is not None:
        out["clocks"] = clocks
    if not args.no_cpu and not sub and world == 1:
        rates = []
        for m, sc in zip(MIXED, scs[0]):
            r, n, dt = cpu_rate(dict(w, model=m), sc, 6, 1)
            rates.append((m, r, n, dt))
        harm = len(rates) / sum(1.0 / r for _, r, _, _ in rates)
        out["cpu_baseline"] = {"value": harm, "unit": "control-steps/s", "cores": 1, "kind": "port", "host_cores_available": os.cpu_count(),
                               "sample": "; ".join(f"{m}: {n} agents in {dt:.1f} s" for m, _, n, dt in rates) + " (oracle port; harmonic mean over the 1/3-1/3-1/3 mix)"}
    return out


# --------------------------------------------------------------------------------------- reference arm
def _backup_cpu_range(args):
    """one process of the Backup-CBF CPU leg: oracle/backup_cbf.py (the restatement of backup_cbf_qp.py:563-794) on agents [lo, hi)"""
    lo, hi, X, Ur, MOV = args
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import backup_cbf as B
    sc = B.EvadeScene()
    t0 = time.perf_counter()
    for a in range(lo, hi):
        B.solve(sc, X[a], Ur[a], MOV[a])
    return hi - lo, time.perf_counter() - t0


def backup_cpu_rate(X, Ur, MOV, procs, per_proc):
    import multiprocessing as mp
    n = procs * per_proc
    with mp.get_context("spawn").Pool(procs) as pool:
        pool.map(_backup_cpu_range, [(0, 1, X, Ur, MOV)] * procs)            # imports / warm-up
        t0 = time.perf_counter()
        pool.map(_backup_cpu_range, [(k * per_proc, (k + 1) * per_proc, X[:n], Ur[:n], MOV[:n]) for k in range(procs)])
        dt = time.perf_counter() - t0
    return n / dt, n


def run_backup(args, w, cx, steps, warmup, sub=False):
    """Backup-CBF QP: one step = scb_backupcbf_solve over N agents (rollout launch + QP launch)."""
    torch = cx.torch
    from safe_control_b200 import BatchedBackupCBF, HostContext, scenes
    from safe_control_b200.backup import host_solve
    N = w["N"]
    P = 2
    batches = [scenes.make_evade_batch(N, seed=1234 + 101 * q) for q in range(P)]
    dins = [[torch.from_numpy(v).to(cx.dev) for v in b] for b in batches]
    ctrl = BatchedBackupCBF()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for k in range(max(warmup, 3)):
        out = ctrl.solve(*dins[k % P])
    cx.barrier()
    sampler = ClockSampler(cx.local_rank)
    if cx.rank == 0 and not sub:
        sampler.start()
    l0 = ctrl.launches
    e0, e1 = ev(), ev()
    e0.record()
    for k in range(steps):
        out = ctrl.solve(*dins[k % P])
    e1.record(); cx.barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = ctrl.launches - l0
    clocks = sampler.stop() if (cx.rank == 0 and not sub) else None
    st = out["status"]
    # end to end: pageable numpy batch -> H2D + 2 launches + D2H of U / status / intervene / h_min inside scb_backupcbf_solve_host
    hc = HostContext(cx.local_rank)
    host_solve(hc, ctrl.params, *batches[0])
    e_steps = max(3, min(steps, 10))
    t0 = time.perf_counter()
    for k in range(e_steps):
        host_solve(hc, ctrl.params, *batches[k % P])
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
    # one agent alone: the latency a single-robot caller of the drop-in class sees on the device
    one = [v[:1].contiguous() for v in dins[0]]
    for _ in range(3):
        ctrl.solve(*one)
    a, b = ev(), ev()
    a.record()
    for _ in range(20):
        ctrl.solve(*one)
    b.record(); torch.cuda.synchronize()
    ms_max, e2e_max = cx.max_over_ranks([ms, e2e_ms])
    if cx.rank != 0:
        return None
    fl = _profile_json(["r2_flops_backup.json"]).get("flops_per_agent")
    h2d = sum(v.nbytes for v in batches[0]); d2h = N * (16 + 4 + 4 + 8)
    rec = {
        "metric": "control-steps/sec (batched QP solves/s)", "value": cx.world * N / (ms_max * 1e-3), "unit": "control-steps/s",
        "n_gpus": cx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {w['desc']}", "agents_per_gpu": N, "backup_steps": 120, "rows_per_qp": 124,
                   "scene": "scenes.make_evade_batch seed 1234 (+101 per batch): evade hallway + pocket, one bullet per agent",
                   "launch": "2 launches per step (backup_rollout_kernel<5>: 5 lanes per agent, 6 agents per warp; backup_qp_kernel<8,16>), eager",
                   "l2_policy": f"{P} distinct batches alternated; compute-bound (189 MB of rows per step stream through L2)",
                   "solver": {"optimal_frac": float((st == 0).float().mean()), "intervene_frac": float(out["intervene"].float().mean())},
                   "single_agent_latency_ms": a.elapsed_time(b) / 20},
        "e2e": {"value": cx.world * N / (e2e_max * 1e-3), "unit": "control-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps_timed": e_steps, "how": "scb_backupcbf_solve_host on pageable numpy buffers (staged H2D, 2 launches, D2H, sync), wall clock"},
        "gpu_launches": launches,
        "roofline": fp64_roofline(fl * N if fl else None, ms_max,
                                  {"kernel_ms_how": "both launches of one step (rollout ~88 %, QP ~12 %: profiles/r2_launches_backup.csv), CUDA events",
                                   "flops_source": "ncu-counted 2*DFMA + DADD + DMUL per agent (profiles/r2_flops_backup.json) x agents per step" if fl else
                                                   "profiles/r2_flops_backup.json missing",
                                   "note": "chain of dependent fp64 sqrt / div per backup step: issue / FP64-pipe bound (ncu: issue slots 70 %, FP64 pipe 41 % busy), HBM traffic negligible"}),
    }
    if clocks is not None:
        rec["clocks"] = clocks
    if not args.no_cpu and not sub:
        procs = min(16, os.cpu_count() or 1)
        v, n = backup_cpu_rate(*batches[0], procs, 12)
        rec["cpu_baseline"] = {"value": v, "unit": "control-steps/s", "cores": procs, "kind": "port",
                               "sample": f"{n} agents of the same batch through oracle/backup_cbf.py (numpy restatement of backup_cbf_qp.py:563-794, exact QP), {procs} processes"}
    return rec


def _shield_cpu_range(args):
    """one process of the shield CPU leg: oracle/shielding.py (restatement of gatekeeper.py:553-672 / mps.py:59-160), first
    control step of agents [lo, hi) (initial backup commitment + the full backward search)"""
    lo, hi, mode, X, NOMX, NOMU, MOV, STAT = args
    from oracle import backup_cbf as B, shielding as S
    sc = B.EvadeScene()
    t0 = time.perf_counter()
    for a in range(lo, hi):
        S.OracleShield(sc, mode=mode).solve(X[a], NOMX[a], NOMU[a], MOV[a], STAT[a])
    return hi - lo, time.perf_counter() - t0


def shield_cpu_rate(mode, arrays, procs, per_proc):
    import multiprocessing as mp
    n = procs * per_proc
    arrs = [a[:n] for a in arrays]
    with mp.get_context("spawn").Pool(procs) as pool:
        pool.map(_shield_cpu_range, [(0, 1, mode, *arrs)] * procs)
        t0 = time.perf_counter()
        pool.map(_shield_cpu_range, [(k * per_proc, (k + 1) * per_proc, mode, *arrs) for k in range(procs)])
        dt = time.perf_counter() - t0
    return n / dt, n


def run_shield(args, w, cx, steps, warmup, sub=False):
    """gatekeeper / MPS: one step = scb_shield_step over N agents (one launch), shield state resident in HBM;
    every step all agents re-plan from the same inputs (the steady state of the example: an event every control step)."""
    torch = cx.torch
    from safe_control_b200 import BatchedShield, scenes
    N, T, mode = w["N"], 100, w["controller"]
    P = 2
    batches = []
    for q in range(P):
        X, _, MOV = scenes.make_evade_batch(N, seed=1234 + 101 * q)
        NOMX, NOMU = scenes.make_evade_plans(X, T)
        batches.append((X, NOMX, NOMU, MOV, scenes.evade_static_rects(MOV)))
    dins = [[torch.from_numpy(v).to(cx.dev) for v in b] for b in batches]
    sh = BatchedShield(N, mode, None, 0.05, None, T, device=cx.dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for k in range(max(warmup, 3)):
        out = sh.step(*dins[k % P])
    cx.barrier()
    l0 = sh.launches
    e0, e1 = ev(), ev()
    e0.record()
    for k in range(steps):
        out = sh.step(*dins[k % P])
    e1.record(); cx.barrier()
    ms = e0.elapsed_time(e1) / steps
    # end to end: pinned host batch -> device, one launch, inputs + backup flags back
    hin = [[torch.from_numpy(v).pin_memory() for v in b] for b in batches]
    hU = torch.empty((N, 2), dtype=torch.float64).pin_memory(); hB = torch.empty((N,), dtype=torch.int32).pin_memory()
    def estep(k):
        for dst, src in zip(dins[k % P], hin[k % P]):
            dst.copy_(src, non_blocking=True)
        o = sh.step(*dins[k % P])
        hU.copy_(o["U"], non_blocking=True); hB.copy_(o["using_backup"], non_blocking=True)
        torch.cuda.synchronize()
    estep(0)
    e_steps = max(3, min(steps, 10))
    t0 = time.perf_counter()
    for k in range(e_steps):
        estep(k)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
    ms_max, e2e_max = cx.max_over_ranks([ms, e2e_ms])
    if cx.rank != 0:
        return None
    fl = (_profile_json(["r2_flops_shield.json"]).get(mode) or {}).get("flops_per_agent_step")
    h2d = sum(v.nbytes for v in batches[0]); d2h = N * (16 + 4)
    rec = {
        "metric": "control-steps/sec (batched QP solves/s)", "value": cx.world * N / (ms_max * 1e-3), "unit": "control-steps/s",
        "n_gpus": cx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {w['desc']}", "agents_per_gpu": N, "nominal_steps": T, "backup_steps": 120,
                   "scene": "scenes.make_evade_batch seed 1234 (+101 per batch) + make_evade_plans (PD nominal plans)",
                   "launch": "gatekeeper: 2 launches per step (candidate 0 with a thread per agent, the other candidates with 8 lanes per queued agent); MPS: 1 launch (a thread per agent); eager",
                   "l2_policy": f"{P} distinct batches alternated ({h2d / 1e6:.0f} MB of plans + state each, > L2)",
                   "outcome": {"nominal_leg_committed_frac": float((sh.nsteps > 0).float().mean()),
                               "mean_committed_nominal_steps": float(sh.nsteps.float().mean()),
                               "using_backup_frac": float(out["using_backup"].float().mean())}},
        "e2e": {"value": cx.world * N / (e2e_max * 1e-3), "unit": "control-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps_timed": e_steps, "how": "pinned host plans / states / obstacles -> H2D -> scb_shield_step -> D2H of U + backup flags -> sync, wall clock"},
        "gpu_launches": sh.launches - l0,
        "roofline": fp64_roofline(fl * N if fl else None, ms_max,
                                  {"kernel_ms_how": "all launches of one step, CUDA events",
                                   "flops_source": "ncu-counted 2*DFMA + DADD + DMUL per agent-step on the tools/prof_shield.py scene (profiles/r2_flops_shield.json; "
                                                   "the count depends on how many candidates an agent has to try) x agents per step" if fl else "profiles/r2_flops_shield.json missing",
                                   "hbm_algorithmic_bytes_per_step": h2d,
                                   "note": "rollout-bound: every candidate is a chain of 120 dependent closed-loop steps (fp64 sqrt / div); issue slots 70 % busy at "
                                           "8-13 of 32 threads active per instruction (profiles/r2_ncu_shield_summary.txt, r2_launches_shield.csv)"}),
    }
    if not args.no_cpu and not sub:
        procs = min(16, os.cpu_count() or 1)
        v, n = shield_cpu_rate(mode, batches[0], procs, 4 if mode == "gatekeeper" else 16)
        rec["cpu_baseline"] = {"value": v, "unit": "control-steps/s", "cores": procs, "kind": "port",
                               "sample": f"first control step of {n} agents of the same batch through oracle/shielding.py, {procs} processes"}
    return rec


def reference_arm(args, w):
    """The reference's own CPU implementation of this path (cvxpy->GUROBI, do-mpc->IPOPT) cannot be installed (no
    network, no wheels); this arm times the reference-equivalent CPU path: the oracle port, one agent at a time per
    process like tracking.py:control_step, on every host core.  One "step" = a bounded sample of the workload
    (per_step agents); value = agent-steps/s over the K timed steps."""
    from safe_control_b200 import scenes
    procs = os.cpu_count() or 1
    if w.get("loop"):
        return reference_loop(args, w, procs)
    if w["controller"] in ("gatekeeper", "mps"):
        X, _, MOV = scenes.make_evade_batch(1024, seed=1234)
        NOMX, NOMU = scenes.make_evade_plans(X, 100)
        val, n = shield_cpu_rate(w["controller"], (X, NOMX, NOMU, MOV, scenes.evade_static_rects(MOV)), procs, 4 if w["controller"] == "gatekeeper" else 16)
        sample = f"first control step of {n} agents of the shield scene (seed 1234) through oracle/shielding.py, {procs} processes"
        print(json.dumps({
            "impl": "reference", "metric": "control-steps/sec (batched QP solves/s)", "value": val, "unit": "control-steps/s",
            "n_gpus": args.gpus, "steps": 1, "warmup": 1, "ms_per_step": 1e3 * n / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": f"{w['name']}: {w['desc']}"},
            "cpu_baseline": {"value": val, "unit": "control-steps/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "control-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "numpy restatement of shielding/gatekeeper.py:553-672 / mps.py:59-160 (the reference itself is pure numpy; its example needs matplotlib)"}))
        return
    if w["controller"] == "backup_cbf_qp":
        batch = scenes.make_evade_batch(4096, seed=1234)
        per = max(4, min(16, 2 * args.steps))
        val, n = backup_cpu_rate(*batch, procs, per)
        sample = f"{n} agents of the backup scene (seed 1234) through oracle/backup_cbf.py, {procs} processes"
        print(json.dumps({
            "impl": "reference", "metric": "control-steps/sec (batched QP solves/s)", "value": val, "unit": "control-steps/s",
            "n_gpus": args.gpus, "steps": 1, "warmup": 1, "ms_per_step": 1e3 * n / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": f"{w['name']}: {w['desc']}"},
            "cpu_baseline": {"value": val, "unit": "control-steps/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "control-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "reference's cvxpy / OSQP stack is not installable here; this is the numpy restatement of backup_cbf_qp.py:563-794"}))
        return
    per_step = {"cbf_qp": 16, "optimal_decay_cbf_qp": 64, "mpc_cbf": 3}[w["controller"]] * procs
    budget_s = 120.0
    slice_s = None
    mixed = w["model"] == "mixed"
    models = list(MIXED) if mixed else [w["model"]]
    n_scene = min(max(per_step * 8, 2048), 16384) if not mixed else max(per_step * 8, 256)
    scs = [scenes.make_scene(m, n_scene, w["M"], seed=1234, dynamic=w["dynamic"],
                             optimal_decay=w["controller"] == "optimal_decay_cbf_qp") for m in models]

    def one(k):
        """one step: per_step agents of EVERY model group (mixed: the 1/3-1/3-1/3 mix) -> (agents, seconds)"""
        n_tot, t_tot = 0, 0.0
        for m, sc in zip(models, scs):
            r, n, dt = cpu_rate(dict(w, model=m), sc, per_step, procs, offset=k * per_step, budget=slice_s)
            n_tot += n; t_tot += dt
        return n_tot, t_tot

    for k in range(max(1, min(args.warmup, 3)) if w["controller"] != "mpc_cbf" else 0):
        one(k)                                   # (MPC: no warm-up step, the oracle has no state to warm and a step costs ~20 s)
    t_tot, n_tot, done = 0.0, 0, 0
    for k in range(args.steps):
        n, dt = one(k + 3)
        t_tot += dt; n_tot += n; done += 1
        if t_tot > budget_s:
            break
    val = n_tot / t_tot
    sample = (f"{done} steps x up to {per_step} agents" + (" of each model group (du / kb / quad3d)" if mixed else "") +
              (f" ({n_tot} agents solved; agents handed to the workers one at a time, SLSQP capped at 40 iterations)" if w["controller"] == "mpc_cbf" else "") +
              f" of the {w['name']} scene (seed 1234), oracle port (numpy/scipy), "
              f"{procs} processes" + ("" if done == args.steps else f"; stopped at the {budget_s:.0f} s budget"))
    print(json.dumps({
        "impl": "reference", "metric": "control-steps/sec (batched QP solves/s)", "value": val, "unit": "control-steps/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(done, 1),
        "higher_is_better": True, "scaling": "strong" if mixed else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {w['desc']}", "agents_per_step": per_step * len(models), "obstacles": w["M"],
                   "horizon": w["H"]},
        "cpu_baseline": {"value": val, "unit": "control-steps/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "control-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference's cvxpy/GUROBI + do-mpc/IPOPT stack is not installable here (no network, no wheels); this is "
                "the oracle restatement driven one agent at a time like tracking.py:control_step",
    }))
    _close_pools()


# --------------------------------------------------------------------------------------- main
SUB_STEPS = {"cfg3": (10, 3), "cfg4": (200, 10), "cfg5": (2, 1), "backup": (10, 3), "gatekeeper": (10, 3), "mps": (20, 3)}     # (steps, warmup) of the sub-records of the N = 1 line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sub", action="store_true", help="N = 1 default run: skip the cfg3 / cfg4 / cfg5 sub-records")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    default_run = args.workload is None
    # default workload: cfg2 on one GPU (BASELINE configs[1]); with more ranks the config BASELINE names for them, cfg5
    name = args.workload or ("cfg2" if max(world, args.gpus) <= 1 else "cfg5")
    w = dict(WORKLOADS[name], name=name)

    if args.impl == "reference":
        if rank != 0:
            return
        return reference_arm(args, w)

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if w.get("loop"):
        return run_loop(args, w, rank, world, local_rank)
    cx = Ctx()
    steps, warmup = args.steps, max(args.warmup, 3)
    if name == "cfg5":
        if default_run:
            steps = min(steps, 5)                  # ~40 ms x N_gpu^-1 ... 300 ms per step: a handful is plenty
        out = run_cfg5(args, w, cx, steps, min(warmup, 3))
    elif name == "backup":
        out = run_backup(args, w, cx, min(steps, 50), min(warmup, 5))
    elif name in ("gatekeeper", "mps"):
        out = run_shield(args, w, cx, min(steps, 50), min(warmup, 5))
    else:
        out = run_single(args, w, cx, steps, warmup)
        if default_run and world == 1 and not args.no_sub:
            subs = {}
            for sname, (s_steps, s_warm) in SUB_STEPS.items():
                sw = dict(WORKLOADS[sname], name=sname)
                try:
                    rec = run_cfg5(args, sw, cx, s_steps, s_warm, sub=True) if sname == "cfg5" else \
                        (run_backup(args, sw, cx, s_steps, s_warm, sub=True) if sname == "backup" else
                         (run_shield(args, sw, cx, s_steps, s_warm, sub=True) if sname in ("gatekeeper", "mps") else
                          run_single(args, sw, cx, s_steps, s_warm, sub=True)))
                    subs[sname] = {k: rec[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "config", "e2e",
                                                       "gpu_launches", "roofline") if k in rec}
                except Exception as e:             # a sub-record must never take the headline down with it
                    subs[sname] = {"error": f"{type(e).__name__}: {e}"}
            out["sub_records"] = subs
    if rank == 0 and out is not None:
        print(json.dumps(out))
    cx.close()


if __name__ == "__main__":
    main()
