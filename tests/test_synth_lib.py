"""The synthetic-world generator is its own library (synth/libbos_synth.so), not part of the product library: the CPU reference
arm of bench.py builds its workload without mapping libbos_b200.so."""
import ctypes
import subprocess
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_library_exports_and_is_deterministic():
    from synth import synth
    L = ctypes.CDLL(synth.build())
    for s in ("bos_synth_default_spec", "bos_synth_create", "bos_synth_destroy", "bos_synth_counts", "bos_synth_get"):
        assert hasattr(L, s)
    a = synth.synth_world(500, 100, 5000, seed=7)
    b = synth.synth_world(500, 100, 5000, seed=7)
    c = synth.synth_world(500, 100, 5000, seed=8)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert not np.array_equal(a["b_z"], c["b_z"])
    # values are float32-exact like the g2o loader's std::stof (utils/g2o_utils.cpp)
    assert np.array_equal(a["b_z"], a["b_z"].astype(np.float32).astype(np.float64))
    assert len(np.unique(np.stack([a["b_pose_id"], a["b_lm_id"]]), axis=1).T) == len(a["b_z"])       # no duplicate (pose, lm) pairs
    assert np.bincount(np.searchsorted(a["lm_ids"], a["b_lm_id"]), minlength=100).min() >= 2          # every landmark seen twice


def test_product_library_has_no_synth_symbols_and_the_generator_does_not_map_it():
    from prb_project_bearing_only_slam_b200 import capi
    L = ctypes.CDLL(capi.LIB_PATH)
    assert not hasattr(L, "bos_synth_create")
    code = ("import sys; sys.path.insert(0, %r); import bench; w, pr, s = bench.make_world('synth-20k'); "
            "maps = open('/proc/self/maps').read(); assert 'libbos_synth' in maps and 'libbos_b200' not in maps, maps; print(pr.Eb)" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
