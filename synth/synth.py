"""ctypes front end of synth/libbos_synth.so (bos_synth.h): the synthetic-world generator of bench.py and the tests.

Kept out of libbos_b200.so on purpose: the CPU reference arm of the benchmark builds its workload with this module and never
maps the product library."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbos_synth.so")
_lib = None


class SynthSpec(C.Structure):
    _fields_ = [("n_poses", C.c_int), ("n_landmarks", C.c_int), ("target_bearing_edges", C.c_int64), ("seed", C.c_uint64),
                ("bearing_sigma", C.c_double), ("odom_sigma_xy", C.c_double), ("odom_sigma_theta", C.c_double),
                ("init_drift", C.c_double), ("init_noise", C.c_double), ("reserved", C.c_int * 8)]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("bos_synth.cpp", "bos_synth.h")]
    stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-shared", "-o", LIB_PATH, srcs[0]])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.bos_synth_default_spec.argtypes = [C.POINTER(SynthSpec)]
        L.bos_synth_default_spec.restype = None
        L.bos_synth_create.argtypes = [C.POINTER(SynthSpec), C.POINTER(vp)]
        L.bos_synth_destroy.argtypes = [vp]
        L.bos_synth_counts.argtypes = [vp, vp]
        L.bos_synth_get.argtypes = [vp] + [vp] * 12
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def synth_world(n_poses, n_landmarks, target_bearing_edges, seed=0xB0500000, **kw):
    """Generates a synthetic world and returns its arrays (ids, initial / true states, measurements)."""
    L = lib()
    spec = SynthSpec()
    L.bos_synth_default_spec(C.byref(spec))
    spec.n_poses, spec.n_landmarks, spec.target_bearing_edges, spec.seed = int(n_poses), int(n_landmarks), int(target_bearing_edges), int(seed)
    for k, v in kw.items():
        setattr(spec, k, v)
    h = C.c_void_p()
    rc = L.bos_synth_create(C.byref(spec), C.byref(h))
    if rc != 0:
        raise ValueError("bos_synth_create: invalid specification")
    try:
        cnt = np.zeros(4, np.int64)
        L.bos_synth_counts(h, _ptr(cnt))
        NP, NL, Eb, Eo = [int(x) for x in cnt]
        w = dict(pose_ids=np.zeros(NP, np.int32), poses_init=np.zeros((NP, 3)), poses_true=np.zeros((NP, 3)),
                 lm_ids=np.zeros(NL, np.int32), lms_true=np.zeros((NL, 2)),
                 b_pose_id=np.zeros(Eb, np.int32), b_lm_id=np.zeros(Eb, np.int32), b_z=np.zeros(Eb),
                 o_src_id=np.zeros(Eo, np.int32), o_dst_id=np.zeros(Eo, np.int32), o_z=np.zeros((Eo, 3)), o_omega=np.zeros((Eo, 9)))
        L.bos_synth_get(h, _ptr(w["pose_ids"]), _ptr(w["poses_init"]), _ptr(w["poses_true"]), _ptr(w["lm_ids"]), _ptr(w["lms_true"]),
                        _ptr(w["b_pose_id"]), _ptr(w["b_lm_id"]), _ptr(w["b_z"]), _ptr(w["o_src_id"]), _ptr(w["o_dst_id"]),
                        _ptr(w["o_z"]), _ptr(w["o_omega"]))
    finally:
        L.bos_synth_destroy(h)
    return w
