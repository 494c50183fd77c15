#!/usr/bin/env python
"""Benchmark of the hot path: full Gauss-Newton iterations (Solver::step) on a synthetic bearing-only world.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on the host cores: complete GN steps with a
                                                             # sparse direct solve, run to completion (no extrapolation)
    torchrun ... bench.py --gpus N ...                       # N ranks: edge-sharded linearization + NCCL combine
    python bench.py --workload full|mini|synth-100k|synth-20k|batch-4096   # the other BASELINE.json configs

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
    # name: (poses, landmarks, target bearing edges, solver)   -- BASELINE.json configs[3] (the metric's config), configs[2]
    "synth-2M": (200000, 50000, 2000000, "pcg"),
    "synth-100k": (10000, 2000, 100000, "auto"),
    "synth-20k": (2000, 400, 20000, "auto"),
    # BASELINE.json configs[1], configs[0]: the reference's bundled datasets (tests/golden/*.npz = data/*.g2o parsed into arrays)
    "full": (301, 141, 2132, "auto"),
    "mini": (3, 6, 15, "auto"),
    # BASELINE.json configs[4]: 4096 independent mini-sized problems, one GN iteration per problem per launch
    "batch-4096": (3, 6, 15, "batch"),
}
SEED = 0xB0500003
CPU_STEP_BUDGET_S = 420.0      # the reference arm stops adding steps once it has run this long (a synth-2M step is ~4 minutes of CPU)


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
    """SM clock and throttle reasons under the timed load (B200_PROFILING.md recipe): an `nvidia-smi -lms 100` child process, started before
    the timed region, every line stamped on arrival.  The timed region of the headline workload is ~85 ms long, about one sampling period, so
    the caller keeps the SAME load running (untimed steps of the same workload, `keep_load`) after its timing is closed until three samples
    have arrived; summary(t0, t1) says how many of them fell inside the timed region itself.  (Polling NVML from a thread of this process
    was tried: its queries contend with the kernel launches and slowed the timed steps by 2 %.)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [int(x) for x in vis.split(",") if x.strip() != ""]
            index = ids[index] if ids else index
        except (ValueError, IndexError):
            pass
        self.index = index
        self.rows = []          # (arrival time, sm_mhz, sm_max_mhz, reasons)
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
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]),
                                  [nm for k, nm in enumerate(self.NAMES) if r[4 + k].lower().startswith("active")]))
            except Exception:
                pass

    def keep_load(self, one_more_step, want=3, limit_s=2.0, fixed=None):
        """Untimed steps of the same workload until `want` samples have arrived (or the child process is gone / limit_s has passed).
        `fixed`: exactly that many steps instead (several ranks: the steps hold collectives, every rank must take the same number)."""
        t = time.perf_counter()
        extra = 0
        if fixed is not None:
            for _ in range(fixed):
                one_more_step()
            extra = fixed
        else:
            while self.proc and self.proc.poll() is None and len(self.rows) < want and time.perf_counter() - t < limit_s:
                one_more_step()
                extra += 1
        self.extra_steps = extra

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        inside = sum(1 for r in rows if t0 is not None and t0 <= r[0] <= t1)
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
                "reasons": sorted({nm for r in rows for nm in r[3]}), "samples": len(rows), "samples_inside_timed_region": inside,
                "sampling": "nvidia-smi -lms 100 from before the timed region; the same load kept running (%d untimed steps) until %d samples"
                            % (getattr(self, "extra_steps", 0), len(rows))}


def make_world(name):
    from prb_project_bearing_only_slam_b200.problem import Problem          # pure Python (id -> stix bookkeeping)
    NP, NL, E, solver = WORKLOADS[name]
    if name in ("full", "mini", "batch-4096"):
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", ("mini" if name == "batch-4096" else name) + ".npz")))
        w = dict(pose_ids=g["pose_ids"], poses_init=g["poses_xyt"], b_pose_id=g["b_pose_id"], b_lm_id=g["b_lm_id"], b_z=g["b_z"],
                 o_src_id=g["o_src_id"], o_dst_id=g["o_dst_id"], o_z=g["o_z"], o_omega=g["o_omega"], lms_tri=g["lms_tri_f64"])
        pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
                     fixed_pose_id=int(g["fixed_pose_id"]))
        return w, pr, solver
    from synth import synth_world                                           # own library: the CPU arm never maps libbos_b200.so
    w = synth_world(NP, NL, E, seed=SEED)
    pr = Problem(w["pose_ids"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"], w["o_omega"],
                 fixed_pose_id=int(w["pose_ids"][0]))
    return w, pr, solver


def data_label(name):
    return "synthetic" if name.startswith("synth") else ("bundled dataset of the reference (data/*.g2o as arrays in tests/golden/)" if name in ("full", "mini")
                                                          else "synthetic (4096 perturbed copies of the bundled mini problem)")


# ---- CPU arm: the reference's algorithm on the host cores ---------------------------------------------------------------
def cpu_oracle(w, pr):
    from oracle.oracle import Oracle
    o = Oracle("f64")
    o.set_problem(w["pose_ids"], w["poses_init"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"],
                  w["o_omega"], fixed_id=pr.fixed_pose_id)
    o.triangulate()
    o.solver_init(pr.fixed_pose_id)
    return o


def cpu_step_superlu(o, pr):
    """One COMPLETE GN step on the CPU with the sparse direct solve done by SuperLU (scipy.sparse.linalg.splu, symmetric mode,
    MMD ordering of A^T + A, no pivoting): linearize + assemble (oracle, O(E)), export H_nofixed, factorise, solve, boxplus."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    t = {}
    t0 = time.perf_counter(); o.linearize(); t["linearize"] = time.perf_counter() - t0
    t0 = time.perf_counter(); colptr, rowidx, val, b = o.csc(); n = len(b)
    A = sp.csc_matrix((val, rowidx, colptr), shape=(n, n)); t["export"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    lu = spl.splu(A, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    t["factor"] = time.perf_counter() - t0
    t0 = time.perf_counter(); x = lu.solve(-b); t["trisolve"] = time.perf_counter() - t0
    t["residual"] = float(np.abs(A @ x + b).max() / max(np.abs(b).max(), 1e-300))
    t["nnzL"] = int(lu.L.nnz)
    d = np.zeros(3 * pr.NP + 2 * pr.NL)
    keep = np.ones(len(d), bool); keep[3 * pr.fixed_stix:3 * pr.fixed_stix + 3] = False
    d[keep] = x
    t0 = time.perf_counter(); o.set_delta(d); o.apply_boxplus(); t["update"] = time.perf_counter() - t0
    t["step"] = t["linearize"] + t["export"] + t["factor"] + t["trisolve"] + t["update"]
    return t


def cpu_step_ldlt(o, deadline_s=0.0):
    """One GN step with the oracle's own sparse LDL^T = the reference's Eigen::SimplicialLDLT restated (slam/solver.hpp:72,
    solver.cpp:75-94): minimum-degree ordering + symbolic phase on the first call only (analyzePattern once), scalar up-looking
    numeric factorisation + solve per step.  With a deadline the factorisation may stop early: `finished` False, `flops_done` says
    how far it got (exact counts from the symbolic phase)."""
    t = {}
    t0 = time.perf_counter(); o.linearize(); t["linearize"] = time.perf_counter() - t0
    info = o.solve_sparse(deadline_s)
    t.update(export=info["t_export"], order=info["t_order"], analyze=info["t_analyze"], factor=info["t_factor"], trisolve=info["t_trisolve"],
             nnzL=info["nnzL"], flops=info["flops"], flops_done=info["flops_done"], finished=info["finished"])
    if info["finished"]:
        t0 = time.perf_counter(); o.apply_boxplus(); t["update"] = time.perf_counter() - t0
        t["step"] = t["linearize"] + t["export"] + t["factor"] + t["trisolve"] + t["update"]
    return t


def cpu_baseline_sample(w, pr, budget_s=20.0):
    """`cpu_baseline` of our arm: a BOUNDED sample of one reference-algorithm GN step on the same workload, single thread (the
    reference has no threads: CMakeLists.txt:5).  The H, b build and the symbolic phase run in full; the numeric LDL^T runs for at
    most `budget_s` seconds and, if it did not finish, its time is scaled by flops_total / flops_done (exact counts from the
    symbolic phase -- the complete, unscaled measurement is what `--impl reference` prints)."""
    o = cpu_oracle(w, pr)
    t_lin = o.time_linearize(3)
    r = cpu_step_ldlt(o, deadline_s=budget_s)
    scale = 1.0 if r["finished"] else r["flops"] / max(r["flops_done"], 1.0)
    t_factor = r["factor"] * scale
    t_tri = r["trisolve"] if r["finished"] else 2.5 * 8 * r["nnzL"] / 2.0e9     # two sweeps over L (12 B per entry) at ~2 GB/s/sweep-equivalent
    t_step = t_lin + r["export"] + t_factor + t_tri
    E = pr.Eb + pr.Eo
    what = ("complete" if r["finished"] else "numeric factorisation stopped after %.0f s at %.1f %% of its %.3g flops and scaled" % (r["factor"], 100.0 / scale, r["flops"]))
    return {"value": 1.0 / t_step, "unit": "iterations/s", "cores": 1, "kind": "port", "edges_linearized_per_s": E / t_lin,
            "phases_s": {"linearize": t_lin, "export_csc": r["export"], "factor": t_factor, "trisolve": t_tri,
                         "order_once": r["order"], "analyze_once": r["analyze"]},
            "nnzL": r["nnzL"], "factor_flops": r["flops"],
            "sample": "CPU oracle = the reference's algorithm restated (it needs Eigen3 / OpenCV, absent here), 1 thread, %s workload: H, b build timed in "
                      "full (%.3f s), sparse LDL^T (SimplicialLDLT restated: minimum-degree ordering + symbolic phase once, up-looking numeric "
                      "phase per step): %s" % (pr_name(pr), t_lin, what)}


def pr_name(pr):
    return "%d-pose / %d-landmark / %d-edge" % (pr.NP, pr.NL, pr.Eb + pr.Eo)


def run_reference(args):
    """--impl reference: COMPLETE Gauss-Newton steps of the reference's algorithm on the host cores, nothing extrapolated.
    linearize + assemble = the CPU oracle (O(E) restatement of slam/solver.cpp:31-69); linear solve = a sparse direct factorisation of
    H_nofixed, as the reference does with Eigen::SimplicialLDLT (slam/solver.hpp:72): `ldlt` = the oracle's own restatement of that
    solver (scalar up-looking LDL^T, the same algorithm class and speed), `superlu` = SuperLU through scipy (supernodal: a STRONGER
    CPU solver than the reference's, used where the restatement would not finish inside the driver's time limit: 661 s per
    factorisation at synth-2M, profiles/ref_ldlt_2M_r02.json).  Never maps libbos_b200.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.perf_counter()
    if args.workload == "batch-4096":
        return run_reference_batch(args)
    w, pr, _ = make_world(args.workload)
    E = pr.Eb + pr.Eo
    o = cpu_oracle(w, pr)
    solver = args.ref_solver
    first = None
    if solver in ("auto", "ldlt"):
        # the symbolic phase (once per problem, like analyzePattern) tells the factorisation's flop count exactly
        o.linearize()
        first = cpu_step_ldlt(o, deadline_s=(1e-9 if solver == "auto" else 0.0))
        if solver == "auto":
            solver = "ldlt" if first["flops"] / 0.8e9 < 60.0 else "superlu"      # ~0.8 GFLOP/s: what the scalar up-looking kernel sustains
            first = None
            o = cpu_oracle(w, pr)
            if solver == "ldlt":
                o.linearize(); o.solve_sparse(1e-9)                               # symbolic phase again on the fresh oracle (untimed, once)
    steps, warm = [], 0
    for i in range(args.warmup + args.steps):
        r = cpu_step_superlu(o, pr) if solver == "superlu" else cpu_step_ldlt(o)
        long_step = r["step"] > 20.0
        if i < args.warmup and not long_step:
            warm += 1
            continue
        steps.append(r)
        if time.perf_counter() - t_start + r["step"] > CPU_STEP_BUDGET_S:
            break
    t_step = statistics.mean(x["step"] for x in steps)
    t_lin = statistics.mean(x["linearize"] for x in steps)
    literal = None
    if args.workload in ("full", "mini"):
        # the reference's LITERAL accumulation (slam/solver.cpp:44,60: an N x N sparse temporary merged into H per edge), FP32 like it
        from oracle.oracle import Oracle
        o32 = Oracle("f32")
        o32.set_problem(w["pose_ids"], w["poses_init"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"],
                        w["o_omega"], fixed_id=pr.fixed_pose_id)
        o32.triangulate(); o32.solver_init(pr.fixed_pose_id)
        t_lit = o32.time_linearize_literal(3 if args.workload == "full" else 200)
        literal = {"linearize_literal_s": t_lit, "edges_linearized_per_s": E / t_lit, "gn_iterations_per_s": 1.0 / (t_lit + t_step - t_lin),
                   "what": "H += J^T Omega J with the reference's per-edge O(N + nnz H) sparse merge, FP32, then the same sparse LDL^T"}
    value = 1.0 / t_step
    restated = None
    if literal:   # bundled datasets: the headline is the reference's LITERAL per-edge accumulation; the O(E) restatement is the secondary
        restated = {"gn_iterations_per_s": value, "edges_linearized_per_s": E / t_lin, "what": "O(E) block-slot assembly (same arithmetic), FP64"}
        t_step = literal["linearize_literal_s"] + t_step - t_lin
        t_lin = literal["linearize_literal_s"]
        value = 1.0 / t_step
    phases = {k: statistics.mean(x[k] for x in steps) for k in ("linearize", "export", "factor", "trisolve", "update")}
    if literal:
        phases["linearize"] = t_lin
    sample = ("%d complete GN step(s) of the %s workload on 1 thread, nothing extrapolated: linearize + assemble %.3f s (CPU oracle, O(E)), sparse direct solve "
              "%.3f s (%s, nnz(L) = %d)%s" %
              (len(steps), args.workload, t_lin, phases["factor"] + phases["trisolve"] + phases["export"],
               "SuperLU via scipy.sparse.linalg.splu, symmetric mode, MMD ordering: supernodal, stronger than the reference's SimplicialLDLT"
               if solver == "superlu" else "the oracle's restatement of Eigen::SimplicialLDLT, symbolic phase cached",
               steps[-1]["nnzL"], "; %d warm-up step(s)" % warm if warm else "; no warm-up (a CPU step of this size is minutes long)"))
    line = {
        "impl": "reference", "metric": "gn_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": len(steps), "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": data_label(args.workload),
        "config": {"workload": args.workload, "poses": pr.NP, "landmarks": pr.NL, "bearing_edges": pr.Eb, "odometry_edges": pr.Eo,
                   "solver": "sparse direct (%s)" % solver},
        "edges_linearized_per_s": E / t_lin,
        "phases_s": phases,
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": 1, "kind": "port", "sample": sample,
                         "edges_linearized_per_s": E / t_lin},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_start,
    }
    if solver == "superlu":
        line["solve_residual"] = steps[-1]["residual"]
    if literal:
        # the reference's OWN compiled sources (oracle/_ref: built against the stand-in for Eigen, whose std::map-based sparse containers are not
        # Eigen's -- a figure beside the headline, not the headline)
        try:
            from oracle import ref as _ref
            if _ref.available():
                rf = _ref.Reference()
                rf.set_problem(w["pose_ids"], w["poses_init"], w["b_pose_id"], w["b_lm_id"], w["b_z"], w["o_src_id"], w["o_dst_id"], w["o_z"],
                               w["o_omega"], fixed_id=pr.fixed_pose_id)
                rf.triangulate(); rf.solver_init(pr.fixed_pose_id)
                rf.step()
                n_ref = 3 if args.workload == "full" else 50
                t0 = time.perf_counter()
                for _ in range(n_ref):
                    rf.step()
                t_ref = (time.perf_counter() - t0) / n_ref
                line["reference_own_sources"] = {"gn_iterations_per_s": 1.0 / t_ref, "steps": n_ref, "kind": "reference",
                                                 "what": "proj02::Solver::step() of the reference's unmodified sources (oracle/_ref), float, compiled against "
                                                         "oracle/eigen_standin instead of Eigen3"}
        except OSError as e:
            line["reference_own_sources"] = {"unavailable": str(e)}
        line["literal_reference_accumulation"] = literal
        line["restated_assembly"] = restated
        line["cpu_baseline"]["sample"] += "; linearize + assemble here = the reference's LITERAL per-edge sparse merge (slam/solver.cpp:44,60), FP32, %.3f s" % t_lin
    print(json.dumps(line), file=args.out, flush=True)


def batch_inputs(nprob, seed=0):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "mini.npz")))
    w, pr, _ = make_world("mini")
    rng = np.random.default_rng(seed)
    from prb_project_bearing_only_slam_b200.problem import xyt_to_xycs
    P0 = xyt_to_xycs(g["poses_xyt"]); L0 = g["lms_tri_f64"]
    poses = np.repeat(P0[None], nprob, 0); lms = np.repeat(L0[None], nprob, 0) + rng.normal(size=(nprob,) + L0.shape) * 0.02
    bz = np.repeat(pr.b_z[None], nprob, 0) + rng.normal(size=(nprob, pr.Eb)) * 0.003
    oz = np.repeat(pr.o_z[None], nprob, 0) + rng.normal(size=(nprob,) + pr.o_z.shape) * 0.01
    return g, pr, poses, lms, bz, oz


def run_reference_batch(args):
    """config 5 on the CPU: the oracle steps the same 4096 mini-sized problems one after the other (sparse LDL^T each)."""
    from oracle.oracle import Oracle
    nprob = 4096
    g, pr, poses, lms, bz, oz = batch_inputs(nprob)
    sample_n = 512
    t0 = time.perf_counter()
    for k in range(sample_n):
        ok = Oracle("f64")
        ok.set_problem(g["pose_ids"], g["poses_xyt"], g["b_pose_id"], g["b_lm_id"], bz[k], g["o_src_id"], g["o_dst_id"], oz[k],
                       pr.o_omega, fixed_id=pr.fixed_pose_id, lm_ids=pr.lm_ids, lms_xy=lms[k])
        ok.solver_init(pr.fixed_pose_id)
        ok.set_state(poses[k], lms[k])
        ok.step(2)
    dt = (time.perf_counter() - t0) / sample_n
    value = 1.0 / dt
    line = {"impl": "reference", "metric": "gn_iterations_per_s", "value": value, "unit": "problem-iterations/s", "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": 1e3 * dt * nprob, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": data_label("batch-4096"), "config": {"workload": "batch-4096", "problems": nprob, "poses": pr.NP, "landmarks": pr.NL},
            "cpu_baseline": {"value": value, "unit": "problem-iterations/s", "cores": 1, "kind": "port",
                             "sample": "%d of the 4096 problems stepped one after the other by the CPU oracle (problem set-up included), 1 thread" % sample_n},
            "e2e": {"value": value, "unit": "problem-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
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

    if args.workload == "batch-4096":
        return run_ours_batch(args, torch, dist, rank, world, local)
    w, pr, solver_name = make_world(args.workload)
    if args.solver != "workload":
        solver_name = args.solver
    solver = {"pcg": capi.SOLVER_PCG, "dense": capi.SOLVER_DENSE_CHOLESKY, "auto": capi.SOLVER_AUTO, "sparse": capi.SOLVER_SPARSE_CHOLESKY}[solver_name]
    prec = capi.PRECISION_F64 if args.precision == "f64" else capi.PRECISION_F32
    S = 8 if args.precision == "f64" else 4
    ctx = capi.Context(device=local, solver=solver, precision=prec, pcg_rtol=args.pcg_rtol, pcg_max_iters=args.pcg_max_iters,
                       pcg_precond=args.pcg_precond)
    if args.device_setup is not None:
        ctx.set_device_setup(bool(args.device_setup))
    t_up = time.perf_counter()
    pr.upload(ctx)
    t_up = time.perf_counter() - t_up
    setup_dev_ms, setup_host_ms = ctx.last_setup_ms()
    auto_mode = args.reduce_mode < 0
    if args.reduce_mode < 0:   # PCG workloads on one box: the pull-based combine over NVLink peer memory; NCCL ownership combine if the mappings cannot be opened
        args.reduce_mode = (5 if world <= 8 else 3) if solver == capi.SOLVER_PCG else 1
    if solver == capi.SOLVER_AUTO:   # what AUTO resolves to is only known after the first solve; the roofline block follows solver_used
        solver = None
    if world > 1:
        uid = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
        if args.reduce_mode >= 4:      # combine over NVLink peer memory: the ranks map each other's value buffers (CUDA IPC)
            ok = 1.0
            try:
                ctx.peer_connect(dist)
            except capi.BosError as e:
                ok = 0.0
                print("rank %d: peer mappings unavailable (%s)" % (rank, e), file=sys.stderr)
            t = torch.tensor([ok], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if t.item() < 1.0:
                if not auto_mode:
                    raise SystemExit("--reduce-mode %d needs CUDA IPC peer mappings between the ranks" % args.reduce_mode)
                args.reduce_mode = 3
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
        if world == 1:
            clk.keep_load(ctx.step)
        else:      # ~0.35 s of the same load, the same number of steps on every rank (rank 0 decides)
            nx = torch.tensor([min(400, int(0.35 / max((t1 - t0) / args.steps, 1e-4)) + 1)], device="cuda")
            dist.broadcast(nx, 0)
            clk.keep_load(ctx.step, fixed=int(nx.item()))
    elapsed = t1 - t0
    clk_region = (t0, t1)
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
    clocks = clk.summary(*clk_region)
    traffic = load_traffic()
    pcg_iters = statistics.mean(s["pcg_iterations"] for s in stats)
    if solver is None:
        solver = int(stats[-1]["solver_used"])
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
                    "traffic": (t_it * pcg_iters) if (t_it and world == 1) else None,
                    "traffic_source": "profiles/traffic.json: dram__bytes of one ncu --set full capture of this command at N = 1, scaled to this run's CG iterations (not re-measured here)",
                    "peak_source": peak_src,
                    "bytes_per_launch": b_it * pcg_iters, "bytes_per_cg_iteration": b_it, "cg_iterations_per_launch": pcg_iters,
                    "ms_per_launch": ms_solve, "us_per_cg_iteration": 1e3 * ms_solve / pcg_iters,
                    "note": "ms_per_launch is the solve phase: the persistent kernel plus its per-solve setup kernels (Schur preparation, "
                            "chain factorisation, coarse operator assembly / Cholesky / inverse), CUDA events on the context's stream"}
    elif solver == capi.SOLVER_DENSE_CHOLESKY and 3 * pr.NP >= 2048:
        # the dominant kernels of a dense step are the FP64 tensor-pipe (DMMA) trailing updates of the blocked Cholesky
        n3 = 3.0 * pr.NP
        flops = n3 ** 3 / 3.0
        p64 = os.path.join(ROOT, "profiles", "fp64_peak_r02.json")
        peak64 = float(json.load(open(p64))["fp64_gemm_tflops_sustained"]) if os.path.exists(p64) else 35.3
        ach = flops / (ms_solve * 1e-3) / 1e12
        roofline = {"kernel": "dense Cholesky of the reduced pose system (k_syrk_big: mma.sync.m8n8k4.f64 trailing updates; tcgen05 has no FP64 kind)",
                    "bound": "tensor", "achieved": ach, "peak": peak64, "unit": "TFLOP/s", "frac": ach / peak64, "traffic": None,
                    "peak_source": "measured cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/fp64_peak_r02.json)", "flops_per_launch": flops,
                    "ms_per_launch": ms_solve, "note": "ms_per_launch is the whole solve phase (Schur assembly, factorisation, triangular solves)"}
    else:
        roofline = {"kernel": "H,b build: k_linearize_odometry (+ block init) + k_linearize_bearing_persistent", "bound": "hbm", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get("hb_build_dram_bytes") if world == 1 else None, "peak_source": peak_src,
                    "bytes_per_launch": bytes_build, "ms_per_launch": ms_lin_kernel}
    line = {
        "metric": "gn_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.precision, "data": data_label(args.workload),
        "ms_per_step_device_events": statistics.mean(s["ms_linearize"] + s["ms_allreduce"] + s["ms_solve"] + s["ms_update"] for s in stats),
        "config": {"workload": args.workload, "poses": pr.NP, "landmarks": pr.NL, "bearing_edges": pr.Eb, "odometry_edges": pr.Eo,
                   "N": int(pi.N), "solver": {0: "schur+pcg(block-tridiagonal chain + coarse-space preconditioner)", 1: "schur+block-jacobi-pcg",
                              2: "schur+pcg(block-tridiagonal chain preconditioner)"}[args.pcg_precond]
                   if solver == capi.SOLVER_PCG else ("schur+skyline-cholesky" if solver == capi.SOLVER_SPARSE_CHOLESKY else "schur+dense-cholesky"),
                   "pcg_rtol": args.pcg_rtol, "pcg_coarse": "4 nodes per chunk, inverse kept for 8 solves (the period doubles while rebuilds stop paying)" if solver == capi.SOLVER_PCG and args.pcg_precond == 0 else None,
                   "parallelism": "edge-shard x%d + %s, solve replicated (rank 0's increment broadcast over NCCL)" %
                   (world, {0: "nccl allreduce(full H,b)", 1: "nccl allreduce(b,diag,pose-pose)+allgather(pose-landmark)", 2: "nccl allreduce(b,diag,pose-pose)",
                           3: "nccl ownership: allreduce(landmark blocks, b_l) + gather of the owned pose ranges",
                           4: "ownership, no collective: the bearing kernel stores / adds into every rank's replica over NVLink peer memory",
                           5: "ownership, no collective: local build, then bulk pulls of the owners' pose ranges and rank-ordered sums of the landmark parts over NVLink peer memory"}[args.reduce_mode]) if world > 1 else "single gpu",
                   "l2": "no flush: value + edge buffers (%.0f MB) exceed the 126 MB L2" % ((int(pi.vals_len) * S + pr.Eb * 24) / 1e6)},
        "edges_linearized_per_s": E / (ms_lin * 1e-3),
        "phases_ms": {"linearize": ms_lin_kernel, "allreduce": ms_allreduce, "solve": ms_solve, "update": ms_update},
        "pcg_iterations": pcg_iters,
        "pcg_iterations_per_step": [int(x["pcg_iterations"]) for x in stats],
        "solver_status_per_step": [int(x.get("solver_status", 0)) for x in stats],
        "chi2_last": stats[-1]["chi2_bearing"] + stats[-1]["chi2_odometry"],
        "state_digest_last": stats[-1]["state_digest"],
        "precond_used": sorted(set(int(x["precond_used"]) for x in stats)), "pcg_resolves": int(sum(x["pcg_resolves"] for x in stats)),
        "roofline": roofline,
        "roofline_linearize": {"kernel": "H,b build: k_linearize_odometry (+ block init) + k_linearize_bearing_persistent", "bound": "hbm", "achieved": achieved, "peak": peak,
                               "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get("hb_build_dram_bytes") if world == 1 else None, "peak_source": peak_src,
                               "bytes_per_launch": bytes_build, "ms_per_launch": ms_lin_kernel},
        "e2e": {"value": args.steps / e2e_elapsed, "unit": "iterations/s",
                "h2d_bytes_per_step": int((4 * pr.NP + 2 * pr.NL) * 8), "d2h_bytes_per_step": int((4 * pr.NP + 2 * pr.NL) * 8 + 64)},
        "gpu_launches": int(sum(s["gpu_launches"] for s in stats)),
        "setup_ms": {"upload_problem": 1e3 * t_up, "pattern_core_on_device": setup_dev_ms, "pattern_on_host": setup_host_ms,
                     "device_setup": setup_dev_ms > 0.0},
        "clocks": clocks,
    }
    if world > 1:
        # NCCL combine of the sharded H, b against the NVLink roofline (B200_PROFILING.md: 770 GB/s measured peer copy per direction per GPU,
        # 900 nominal).  Bytes a rank must RECEIVE per step: a ring / tree all-reduce moves 2 (N-1)/N of the summed payload through every
        # GPU's link, a gather brings in the (N-1)/N of the gathered arrays that other ranks own.
        f = (world - 1) / world
        n_pp, Ntot = int(pi.n_hpp_off), int(pi.N)
        summed, gathered = {0: (int(pi.vals_len), 0), 1: (Ntot + 6 * pr.NP + 3 * pr.NL + 9 * n_pp, 6 * pr.Eb),
                            2: (Ntot + 6 * pr.NP + 3 * pr.NL + 9 * n_pp, 0), 3: (5 * pr.NL, 9 * pr.NP), 4: (5 * pr.NL, 9 * pr.NP), 5: (5 * pr.NL, 9 * pr.NP)}[args.reduce_mode]
        nv_peak = 770.0
        if args.reduce_mode == 4:
            # fused: a rank's link carries OUT its owned pose blocks to the N-1 other replicas and (at least) its share of the landmark parts to
            # each of them; there is no separate combine phase, so the time is the whole sharded build (both barriers included)
            nv_bytes = (world - 1) * (gathered / world + summed / world) * S
            t_comb = ms_lin
            what = "k_linearize_bearing_persistent<peer>: stores / REDs into every replica over NVLink + two cross-GPU barriers (reduce_mode 4)"
            note = "fused with the build: time = the whole sharded build, bytes = the lower bound a rank must send (one landmark part per landmark and peer)"
        elif args.reduce_mode == 5:
            # pulled: a rank's link carries IN the pose ranges it does not own and every other rank's landmark parts
            nv_bytes = (f * gathered + (world - 1) * summed) * S
            t_comb = ms_allreduce
            what = "k_peer_pull + k_peer_commit between two cross-GPU barriers (reduce_mode 5)"
            note = "bulk coalesced reads over NVLink peer mappings; the time holds both barriers (launch skew between the ranks included)"
        else:
            nv_bytes = (2 * f * summed + f * gathered) * S
            t_comb = ms_allreduce
            what = "ncclAllReduce / ncclBroadcast of the partial H, b (reduce_mode %d)" % args.reduce_mode
            note = "latency bound at this size: %d collectives of a few MB per step" % (2 + 2 * world if args.reduce_mode == 3 else 1)
        line["roofline_combine"] = {"kernel": what, "bound": "nvlink",
                                    "achieved": nv_bytes / (t_comb * 1e-3) / 1e9, "peak": nv_peak, "unit": "GB/s",
                                    "frac": nv_bytes / (t_comb * 1e-3) / 1e9 / nv_peak, "bytes_per_rank_per_step": int(nv_bytes),
                                    "ms_per_step": t_comb, "summed_scalars": int(summed), "gathered_scalars": int(gathered),
                                    "peak_source": "B200_PROFILING.md: measured peer copy, GB/s per direction per GPU (900 nominal)", "note": note}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(w, pr, budget_s=args.cpu_budget)
    print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ours_batch(args, torch, dist, rank, world, local):
    """BASELINE config 5: 4096 independent mini-sized problems, ONE launch = one GN iteration of every problem (k_batch_step).
    Replicas only across GPUs (every rank steps its own 4096 problems, no collective): value = problem-iterations/s of all ranks."""
    from prb_project_bearing_only_slam_b200 import capi
    nprob = 4096
    g, pr, poses, lms, bz, oz = batch_inputs(nprob, seed=rank)
    prec = capi.PRECISION_F64 if args.precision == "f64" else capi.PRECISION_F32
    S = 8 if args.precision == "f64" else 4
    B = capi.Batch(nprob, pr.NP, pr.NL, pr.fixed_stix, pr.b_pose, pr.b_lm, bz, None, pr.o_src, pr.o_dst, oz, pr.o_omega, precision=prec, device=local)
    B.set_states(poses, lms)
    B.step_device(max(args.warmup, 3))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    with ClockSampler(local) as clk:
        clk_t0 = time.perf_counter()
        ms = B.step_device(args.steps)          # CUDA events around the K launches on the batch's stream
        barrier()
        clk_t1 = time.perf_counter()
        clk.keep_load(lambda: B.step_device(1))
    # end to end: states from pinned host memory, one launch, chi2 / status back, every step
    ph = torch.from_numpy(poses.copy()).pin_memory().numpy(); lh = torch.from_numpy(lms.copy()).pin_memory().numpy()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        B.set_states(ph, lh)
        chi, dinf, st = B.step()
    barrier()
    e2e = time.perf_counter() - t0
    t = torch.tensor([ms, e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    # algorithmic bytes of one launch: per problem the state in and out, the measurements, chi2 / status out (topology is shared)
    b_launch = nprob * ((4 * pr.NP + 2 * pr.NL) * S * 2 + (pr.Eb + 3 * pr.Eo) * S + 32)
    ach = b_launch / (ms / args.steps * 1e-3) / 1e9
    value = world * nprob * args.steps / (ms * 1e-3)
    line = {"metric": "gn_iterations_per_s", "value": value, "unit": "problem-iterations/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": data_label("batch-4096"),
            "config": {"workload": "batch-4096", "problems": nprob, "poses": pr.NP, "landmarks": pr.NL, "bearing_edges": pr.Eb, "odometry_edges": pr.Eo,
                       "parallelism": "replicas only x%d" % world, "l2": "working set %.1f MB stays in L2: launch-latency bound by design" % (b_launch / 1e6)},
            "edges_linearized_per_s": value * (pr.Eb + pr.Eo),
            "roofline": {"kernel": "k_batch_step (one warp per problem, whole GN iteration in shared memory)", "bound": "hbm", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src, "bytes_per_launch": b_launch,
                         "ms_per_launch": ms / args.steps},
            "e2e": {"value": world * nprob * args.steps / e2e, "unit": "problem-iterations/s",
                    "h2d_bytes_per_step": int(nprob * (4 * pr.NP + 2 * pr.NL) * 8), "d2h_bytes_per_step": int(nprob * (2 * 8 + 8 + 4))},
            "gpu_launches": args.steps, "clocks": clk.summary(clk_t0, clk_t1), "status_ok": bool(np.all(st == 0))}
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
    ap.add_argument("--reduce-mode", type=int, default=-1, help="-1: 5 (local build + bulk pulls over NVLink peer memory; 3 = ownership combine over NCCL when the mappings cannot be opened or beyond 8 ranks) for the PCG workloads, 1 for the dense ones; 4 = the bearing kernel pushes into every replica (measured slower, DESIGN.md section 6)")
    ap.add_argument("--ref-solver", default="auto", choices=["auto", "ldlt", "superlu"],
                    help="reference arm: ldlt = the oracle's restatement of Eigen::SimplicialLDLT, superlu = scipy's SuperLU; auto = ldlt when "
                         "its symbolic phase predicts under a minute per factorisation, else superlu")
    ap.add_argument("--solver", default="workload", choices=["workload", "auto", "dense", "pcg", "sparse"])
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds the cpu_baseline sample may spend in the numeric factorisation")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-setup", type=int, default=None, choices=[0, 1], help="build the bearing-edge core of the pattern on the GPU (1) or on host threads (0); default: the library's choice (GPU from 200 k edges on); one-time, outside the timed steps")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 3)
    args.out = _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
