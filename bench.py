#!/usr/bin/env python
"""Benchmark of the hot path: full Gauss-Newton iterations (Solver::step) on a synthetic bearing-only world.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on the host cores (CPU oracle)
    torchrun ... bench.py --gpus N ...                       # N ranks: edge-sharded linearization + NCCL combine

One "step" is one GN iteration: linearize + assemble (+ allreduce) + Schur/PCG solve + boxplus update.
Prints ONE JSON line (rank 0).  `value` = GN iterations/s with the state resident in HBM; `e2e` = the same through
bos_step_host with pinned HOST state buffers (H2D + D2H inside the timed region); `edges_linearized_per_s` and
`roofline` describe the H, b build (K1+K2+K3), timed with CUDA events on the context's own stream inside the
timed steps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (poses, landmarks, target bearing edges, solver)   -- BASELINE.json configs[2], configs[3]
    "synth-2M": (200000, 50000, 2000000, "pcg"),
    "synth-100k": (10000, 2000, 100000, "dense"),
    "synth-20k": (2000, 400, 20000, "dense"),
}
SEED = 0xB0500003


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def hb_build_bytes(NP, NL, Eb, Eo, n_off, S):
    """Algorithmic bytes of one H, b build in OUR layout: SoA edge reads + state once + every block written once.
    reads : bearing 2xint32 + z + omega ; odometry 2xint32 + slot int32 + z[3] + Omega upper[6] ; state 4S/pose + 2S/lm
    writes: pose-lm 3x2 ; pose-pose 3x3 ; diagonal blocks stored symmetric (6 / 3 scalars) ; b."""
    N = 3 * NP + 2 * NL
    reads = Eb * (8 + 2 * S) + Eo * (12 + 9 * S) + NP * 4 * S + NL * 2 * S
    writes = Eb * 6 * S + n_off * 9 * S + NP * 6 * S + NL * 3 * S + N * S
    return reads + writes


def pcg_iteration_bytes(NP, NL, Eb, S):
    """Algorithmic bytes of ONE CG iteration of the persistent PCG kernel: every array the iteration
    must touch, each element once (gathered records counted once per pass, not once per edge):
    landmark pass: 4 B pose word per edge, pose state 4S + z 4S per pose, per landmark Hll^-1 3S + position 2S read, u 2S written;
    pose pass:     2 B landmark-table index per edge, per pose state 4S + z 4S read + z' 4S written + Hpp 6S + M^-1 6S +
                   two pose-pose blocks 12S + two neighbour indices 8 B + x 3S read and written; u, position 4S per landmark."""
    return Eb * 6 + NP * (8 * S + 12 * S + 24 * S + 8 + 6 * S) + NL * (7 * S + 4 * S)


def load_traffic():
    """DRAM bytes per unit from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def recorded_block_jacobi_iterations(workload, default):
    """CG iterations per GN step of the 3x3 block-Jacobi PCG on this workload (measured on the GPU with --pcg-precond 1; the
    CPU restatement runs the same algorithm at ~0.05 s per iteration, so it is extrapolated, not run to convergence)."""
    p = os.path.join(ROOT, "profiles", "pcg_iterations.json")
    try:
        return int(json.load(open(p)).get(workload, default))
    except Exception:
        return default


def make_world(name):
    from synth import synth_world                                           # own library: the CPU arm never maps libbos_b200.so
    from prb_project_bearing_only_slam_b200.problem import Problem          # pure Python (id -> stix bookkeeping)
    NP, NL, E, solver = WORKLOADS[name]
    w = synth_world(NP, NL, E, seed=SEED)
    pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
                 fixed_pose_id=int(w["pose_ids"][0]))
    return w, pr, solver


def oracle_sample(w, pr, pcg_iters_full, pcg_rtol, sample_iters=10):
    """Times the CPU oracle (single thread, like the reference: CMakeLists.txt:5 has no OpenMP) on this workload.
    The H, b build is timed in full; the Schur-PCG solve is timed for 2 and 2+sample_iters CG iterations and
    extrapolated to the iteration count the same algorithm needs at the same tolerance (`pcg_iters_full`)."""
    from oracle.oracle import Oracle
    o = Oracle("f64")
    o.set_problem(w["pose_ids"], w["poses_init"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"],
                  w["o_omega"], fixed_id=pr.fixed_pose_id)
    o.triangulate()
    o.solver_init(pr.fixed_pose_id)
    o.linearize()
    t_lin = o.time_linearize(3)
    t0 = time.perf_counter(); o.solve(1, 2, 0.0); t2 = time.perf_counter() - t0
    t0 = time.perf_counter(); o.solve(1, 2 + sample_iters, 0.0); tk = time.perf_counter() - t0
    t_iter = max((tk - t2) / sample_iters, 1e-9)
    t0 = time.perf_counter(); o.apply_boxplus(); t_upd = time.perf_counter() - t0
    t_step = t_lin + (t2 - 2 * t_iter) + pcg_iters_full * t_iter + t_upd
    return dict(t_lin=t_lin, t_cg_iter=t_iter, t_step=t_step, t_fixed=t2 - 2 * t_iter, t_update=t_upd)


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path (CPU oracle restatement: the reference itself
    needs Eigen3 + OpenCV and cannot be built in this image), single host thread, same workload / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, pr, solver = make_world(args.workload)
    E = pr.Eb + pr.Eo
    iters_full = args.ref_pcg_iters
    times = []
    for i in range(args.warmup + args.steps):
        s = oracle_sample(w, pr, iters_full, args.pcg_rtol, sample_iters=6)
        if i >= args.warmup:
            times.append(s)
    t_step = statistics.mean(x["t_step"] for x in times)
    t_lin = statistics.mean(x["t_lin"] for x in times)
    value = 1.0 / t_step
    sample = ("full %s world; H,b build timed in full (%.3f s); Schur-PCG timed for 2 and 8 CG iterations (%.4f s/iteration) and "
              "extrapolated to %d iterations (what this 3x3 block-Jacobi PCG needs at rtol %.0e: measured on the GPU with --pcg-precond 1); single thread" %
              (args.workload, t_lin, statistics.mean(x["t_cg_iter"] for x in times), iters_full, args.pcg_rtol))
    line = {
        "impl": "reference", "metric": "gn_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "poses": pr.NP, "landmarks": pr.NL, "bearing_edges": pr.Eb, "odometry_edges": pr.Eo,
                   "solver": "schur+block-jacobi-pcg", "pcg_rtol": args.pcg_rtol},
        "edges_linearized_per_s": E / t_lin,
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": 1, "kind": "port", "sample": sample,
                         "edges_linearized_per_s": E / t_lin},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from prb_project_bearing_only_slam_b200 import capi
    from prb_project_bearing_only_slam_b200.problem import xyt_to_xycs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w, pr, solver_name = make_world(args.workload)
    solver = capi.SOLVER_PCG if solver_name == "pcg" else capi.SOLVER_DENSE_CHOLESKY
    prec = capi.PRECISION_F64 if args.precision == "f64" else capi.PRECISION_F32
    S = 8 if args.precision == "f64" else 4
    ctx = capi.Context(device=local, solver=solver, precision=prec, pcg_rtol=args.pcg_rtol, pcg_max_iters=args.pcg_max_iters,
                       pcg_precond=args.pcg_precond)
    pr.upload(ctx)
    if args.reduce_mode < 0:
        args.reduce_mode = 2 if solver == capi.SOLVER_PCG else 1
    if world > 1:
        uid = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
        ctx.set_reduce_mode(args.reduce_mode)
    # initial guess: generated poses + landmarks triangulated ON THE DEVICE (K8)
    P0 = xyt_to_xycs(w["poses_init"])
    ctx.set_state(P0, None)
    ctx.triangulate()
    P0, L0 = ctx.get_state()
    pi = ctx.pattern_info()

    # ---- device-resident arm -------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        ctx.step()
    ctx.set_state(P0, L0)
    stats = []
    barrier()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            stats.append(ctx.step().as_dict())
        barrier()
        t1 = time.perf_counter()
    elapsed = t1 - t0
    # ---- end-to-end arm: host state in, host state out, every step -------------------------------------------------
    Ph = torch.from_numpy(P0.copy()).pin_memory().numpy()
    Lh = torch.from_numpy(L0.copy()).pin_memory().numpy()
    for _ in range(min(args.warmup, 2)):
        ctx.step_host(Ph, Lh)
    Ph[:] = P0; Lh[:] = L0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.step_host(Ph, Lh)
    barrier()
    e2e_elapsed = time.perf_counter() - t0

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- H, b build (+ NCCL combine) on its own, ranks aligned by a barrier before every build: inside a full step the
    # combine's event time also contains the skew the replicated 0.3 s solves accumulate between ranks
    lin_ms, red_ms = [], []
    ctx.set_state(P0, L0)
    for i in range((args.warmup + max(args.steps, 5)) if world > 1 else 0):
        barrier()
        ctx.linearize()
        st_ = ctx.stats()
        if i >= args.warmup:
            lin_ms.append(st_.ms_linearize); red_ms.append(st_.ms_allreduce)
    elapsed = reduce_max(elapsed)
    e2e_elapsed = reduce_max(e2e_elapsed)
    if world == 1:   # single GPU: the build as timed inside the K timed steps
        lin_ms = [s["ms_linearize"] for s in stats]; red_ms = [s["ms_allreduce"] for s in stats]
    ms_lin_kernel = reduce_max(statistics.mean(lin_ms))
    ms_allreduce = reduce_max(statistics.mean(red_ms))
    ms_lin = reduce_max(statistics.mean(a + b for a, b in zip(lin_ms, red_ms)))
    ms_solve = reduce_max(statistics.mean(s["ms_solve"] for s in stats))
    ms_update = reduce_max(statistics.mean(s["ms_update"] for s in stats))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    E = pr.Eb + pr.Eo
    peak, peak_src = load_peaks()
    shard = ctx.edge_shard()
    eb_local, eo_local = shard[1] - shard[0], shard[3] - shard[2]
    # per-launch algorithmic bytes of THIS rank's H, b build (its edge shard; state and diagonal/b prefix in full)
    bytes_build = hb_build_bytes(pr.NP, pr.NL, eb_local, eo_local, int(pi.n_hpp_off), S)
    achieved = bytes_build / (ms_lin_kernel * 1e-3) / 1e9
    value = args.steps / elapsed
    clocks = clk.summary()
    traffic = load_traffic()
    pcg_iters = statistics.mean(s["pcg_iterations"] for s in stats)
    if solver == capi.SOLVER_PCG and pcg_iters > 0:
        # the dominant kernel of a step is the persistent PCG kernel (> 99 % of the step at synth-2M): one launch = one solve
        b_it = pcg_iteration_bytes(pr.NP, pr.NL, pr.Eb, S)
        if args.pcg_precond != 1:   # chain preconditioner: 16 FP32 factor values per pose row are read every CG iteration
            b_it += 16 * 4 * pr.NP
        ach = b_it * pcg_iters / (ms_solve * 1e-3) / 1e9
        t_it = traffic.get("pcg_dram_bytes_per_cg_iteration")
        roofline = {"kernel": "k_pcg_fused (persistent cooperative kernel: the whole %s PCG solve of one GN iteration)" %
                              ({0: "chain + coarse-space preconditioned", 1: "block-Jacobi", 2: "chain-preconditioned"}[args.pcg_precond]),
                    "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": (t_it * pcg_iters) if t_it else None, "peak_source": peak_src,
                    "bytes_per_launch": b_it * pcg_iters, "bytes_per_cg_iteration": b_it, "cg_iterations_per_launch": pcg_iters,
                    "ms_per_launch": ms_solve, "us_per_cg_iteration": 1e3 * ms_solve / pcg_iters,
                    "note": "ms_per_launch is the solve phase: the persistent kernel plus its per-solve setup kernels (Schur preparation, "
                            "chain factorisation, coarse operator assembly / Cholesky / inverse), CUDA events on the context's stream"}
    else:
        roofline = {"kernel": "H,b build: k_landmark_init + k_linearize_bearing_persistent + k_pose_finish", "bound": "hbm", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get("hb_build_dram_bytes"), "peak_source": peak_src,
                    "bytes_per_launch": bytes_build, "ms_per_launch": ms_lin_kernel}
    line = {
        "metric": "gn_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": args.workload, "poses": pr.NP, "landmarks": pr.NL, "bearing_edges": pr.Eb, "odometry_edges": pr.Eo,
                   "N": int(pi.N), "solver": {0: "schur+pcg(block-tridiagonal chain + coarse-space preconditioner)", 1: "schur+block-jacobi-pcg",
                              2: "schur+pcg(block-tridiagonal chain preconditioner)"}[args.pcg_precond]
                   if solver == capi.SOLVER_PCG else "schur+dense-cholesky",
                   "pcg_rtol": args.pcg_rtol, "parallelism": "edge-shard x%d + nccl %s, solve replicated" %
                   (world, {0: "allreduce(full H,b)", 1: "allreduce(b,diag,pose-pose)+allgather(pose-landmark)", 2: "allreduce(b,diag,pose-pose)"}[args.reduce_mode]) if world > 1 else "single gpu",
                   "l2": "no flush: value + edge buffers (%.0f MB) exceed the 126 MB L2" % ((int(pi.vals_len) * S + pr.Eb * 24) / 1e6)},
        "edges_linearized_per_s": E / (ms_lin * 1e-3),
        "phases_ms": {"linearize": ms_lin_kernel, "allreduce": ms_allreduce, "solve": ms_solve, "update": ms_update},
        "pcg_iterations": pcg_iters,
        "pcg_iterations_per_step": [int(x["pcg_iterations"]) for x in stats],
        "solver_status_per_step": [int(x.get("solver_status", 0)) for x in stats],
        "chi2_last": stats[-1]["chi2_bearing"] + stats[-1]["chi2_odometry"],
        "roofline": roofline,
        "roofline_linearize": {"kernel": "H,b build: k_landmark_init + k_linearize_bearing_persistent + k_pose_finish", "bound": "hbm", "achieved": achieved, "peak": peak,
                               "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get("hb_build_dram_bytes"), "peak_source": peak_src,
                               "bytes_per_launch": bytes_build, "ms_per_launch": ms_lin_kernel},
        "e2e": {"value": args.steps / e2e_elapsed, "unit": "iterations/s",
                "h2d_bytes_per_step": int((4 * pr.NP + 2 * pr.NL) * 8), "d2h_bytes_per_step": int((4 * pr.NP + 2 * pr.NL) * 8 + 64)},
        "gpu_launches": int(sum(s["gpu_launches"] for s in stats)),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        # the CPU port runs the reference-arm algorithm (3x3 block-Jacobi PCG): its own iteration count at this tolerance
        cpu_iters = int(round(line["pcg_iterations"]))
        if solver == capi.SOLVER_PCG and args.pcg_precond != 1:
            cpu_iters = recorded_block_jacobi_iterations(args.workload, cpu_iters)
        cb = oracle_sample(w, pr, cpu_iters, args.pcg_rtol)
        line["cpu_baseline"] = {
            "value": 1.0 / cb["t_step"], "unit": "iterations/s", "cores": 1, "kind": "port",
            "edges_linearized_per_s": E / cb["t_lin"],
            "sample": "CPU oracle (restatement; the reference needs Eigen3/OpenCV, absent here), 1 thread, full %s world: H,b build timed "
                      "in full (%.3f s), Schur-PCG timed for 2 and 12 CG iterations (%.4f s/iteration) and extrapolated to the %d "
                      "iterations its 3x3 block-Jacobi preconditioner needs at this tolerance (profiles/pcg_iterations.json)" %
                      (args.workload, cb["t_lin"], cb["t_cg_iter"], cpu_iters)}
    print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout():
    """Libraries (NCCL's version banner, for one) write to file descriptor 1; the contract is ONE JSON line on stdout.  Route fd 1 to
    stderr for the whole run and return a file object on the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synth-2M", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--pcg-rtol", type=float, default=1e-8)
    ap.add_argument("--pcg-max-iters", type=int, default=20000)
    ap.add_argument("--pcg-precond", type=int, default=0, choices=[0, 1, 2],
                    help="0 chain (block-tridiagonal) + coarse-space preconditioner, 1 3x3 block-Jacobi, 2 chain only")
    ap.add_argument("--reduce-mode", type=int, default=-1, help="-1: 2 for the PCG workloads, 1 for the dense ones")
    ap.add_argument("--ref-pcg-iters", type=int, default=0, help="CG iterations per GN step the reference arm extrapolates to "
                    "(0 = the count recorded by our arm in profiles/pcg_iterations.json, else 300)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 3)
    args.out = _quiet_stdout()
    if args.impl == "reference":
        if args.ref_pcg_iters <= 0:
            p = os.path.join(ROOT, "profiles", "pcg_iterations.json")
            args.ref_pcg_iters = int(json.load(open(p)).get(args.workload, 300)) if os.path.exists(p) else 300
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
