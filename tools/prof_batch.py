"""GPU box: BASELINE config 5 -- 4096 independent mini-sized problems, one GN iteration per problem per launch."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden, golden_problem
from prb_project_bearing_only_slam_b200 import capi
from prb_project_bearing_only_slam_b200.problem import xyt_to_xycs

nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = load_golden("mini"); pr = golden_problem(g)
rng = np.random.default_rng(0)
P0 = xyt_to_xycs(g["poses_xyt"]); L0 = g["lms_tri_f64"]
for prec, name in ((capi.PRECISION_F64, "f64"), (capi.PRECISION_F32, "f32")):
    poses = np.repeat(P0[None], nprob, 0); lms = np.repeat(L0[None], nprob, 0) + rng.normal(size=(nprob,) + L0.shape) * 0.02
    bz = np.repeat(pr.b_z[None], nprob, 0) + rng.normal(size=(nprob, pr.Eb)) * 0.003
    oz = np.repeat(pr.o_z[None], nprob, 0) + rng.normal(size=(nprob,) + pr.o_z.shape) * 0.01
    B = capi.Batch(nprob, pr.NP, pr.NL, pr.fixed_stix, pr.b_pose, pr.b_lm, bz, None, pr.o_src, pr.o_dst, oz, pr.o_omega, precision=prec)
    B.set_states(poses, lms)
    B.step_device(3)
    ms = B.step_device(50) / 50
    chi, dinf, st = B.step()
    print("%s: %d problems, %.4f ms per launch (one GN iteration each) -> %.3e problem-iterations/s, %.3e edges/s; status ok %s, chi2 mean %.3e" % (
        name, nprob, ms, nprob / (ms * 1e-3), nprob * (pr.Eb + pr.Eo) / (ms * 1e-3), bool(np.all(st == 0)), chi.sum(1).mean()))
