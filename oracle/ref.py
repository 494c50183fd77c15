"""ctypes front end of oracle/_ref/libbos_ref.so: the REFERENCE'S OWN sources compiled against oracle/eigen_standin.

TEST INFRASTRUCTURE ONLY.  `build()` needs /root/reference (this container); on a box without it only a prebuilt
oracle/_ref/libbos_ref.so can be loaded.  Used by tests/test_ref_build.py, tests/golden/make_ref_golden.py and
bench.py --impl reference (mini / full workloads); never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libbos_ref.so")
REFERENCE_ROOT = os.environ.get("BOS_REFERENCE_ROOT", "/root/reference")
_LIB = None


def sources_present():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "slam", "solver.cpp"))


def build():
    """Compile the reference's sources where they lie (outputs only under oracle/_ref/).  No-op without /root/reference."""
    if sources_present():
        subprocess.check_call(["make", "-C", _HERE, "ref", "REF=" + REFERENCE_ROOT], stdout=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


def available():
    return os.path.exists(_SO) or sources_present()


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise OSError("oracle/_ref/libbos_ref.so is absent and %s is not there to build it from" % REFERENCE_ROOT)
        L = C.CDLL(so)
        L.ref_new.restype = C.c_void_p
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_load_g2o.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_bound.restype = C.c_float
        L.ref_bound.argtypes = [C.c_void_p]
        L.ref_set_params.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.ref_H_nnz.restype = C.c_long
        L.ref_H_nnz.argtypes = [C.c_void_p, C.c_int]
        L.ref_predict_bearing.restype = C.c_float
        L.ref_predict_bearing.argtypes = [C.c_void_p] + [C.c_float] * 5
        L.ref_normalized_angle.restype = C.c_float
        L.ref_normalized_angle.argtypes = [C.c_void_p, C.c_float]
        L.ref_smallest_angle.restype = C.c_float
        L.ref_smallest_angle.argtypes = [C.c_float]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Reference:
    """State + observations + proj02::Solver of the reference, driven through the calls its main loop makes."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.ref_new())

    def __del__(self):
        try:
            self.L.ref_free(self.h)
        except Exception:
            pass

    def load_g2o(self, path):
        if self.L.ref_load_g2o(self.h, path.encode()) != 0:
            raise IOError("reference: cannot open %s" % path)

    def set_problem(self, pose_ids, poses_xyt, b_pose_id, b_lm_id, b_z, o_src_id, o_dst_id, o_z, o_omega,
                    lm_ids=None, lms_xy=None, b_omega=None, fixed_id=-1):
        """Same argument meaning as oracle.Oracle.set_problem; everything is narrowed to float, the reference's scalar."""
        pose_ids = _i(pose_ids); poses_xyt = _f(poses_xyt)
        self.L.ref_add_poses(self.h, len(pose_ids), _p(pose_ids), _p(poses_xyt))
        if lm_ids is not None and len(lm_ids):
            lm_ids = _i(lm_ids); lms_xy = _f(lms_xy)
            self.L.ref_add_landmarks(self.h, len(lm_ids), _p(lm_ids), _p(lms_xy))
        b_pose_id = _i(b_pose_id); b_lm_id = _i(b_lm_id); b_z = _f(b_z)
        bo = None if b_omega is None else _f(b_omega)
        self.L.ref_add_bearings(self.h, len(b_z), _p(b_pose_id), _p(b_lm_id), _p(b_z), _p(bo))
        o_src_id = _i(o_src_id); o_dst_id = _i(o_dst_id); o_z = _f(o_z); o_omega = _f(o_omega)
        self.L.ref_add_odometry(self.h, len(o_src_id), _p(o_src_id), _p(o_dst_id), _p(o_z), _p(o_omega))
        self.L.ref_set_fixed(self.h, int(fixed_id))

    def triangulate(self):
        if self.L.ref_triangulate(self.h) != 0:
            raise KeyError("reference: triangulation hit an unknown pose id")

    def counts(self):
        out = np.zeros(6, np.int32)
        self.L.ref_counts(self.h, _p(out))
        return dict(NP=int(out[0]), NL=int(out[1]), Eb=int(out[2]), Eo=int(out[3]), fixed_pose_id=int(out[4]), N=int(out[5]))

    def bound(self):
        return float(self.L.ref_bound(self.h))

    def ids(self):
        c = self.counts()
        p = np.zeros(c["NP"], np.int32); l = np.zeros(c["NL"], np.int32)
        self.L.ref_get_ids(self.h, _p(p), _p(l))
        return p, l

    def state(self):
        c = self.counts()
        P = np.zeros((c["NP"], 4), np.float32); Lm = np.zeros((c["NL"], 2), np.float32)
        self.L.ref_get_state(self.h, _p(P), _p(Lm))
        return P, Lm

    def state_xyt(self):
        c = self.counts()
        P = np.zeros((c["NP"], 3), np.float32)
        self.L.ref_get_state_xyt(self.h, _p(P))
        return P

    def solver_init(self, fixed_id=-1):
        if self.L.ref_solver_init(self.h, int(fixed_id)) != 0:
            raise KeyError("reference: unknown fixed pose id")

    def set_params(self, kernel_threshold=1.0, damping=0.01):
        self.L.ref_set_params(self.h, kernel_threshold, damping)

    def step(self):
        """proj02::Solver::step().  Returns 0, or 2 when the reference printed its 'not SPD' warning."""
        rc = self.L.ref_step(self.h)
        if rc == 1:
            raise KeyError("reference: step hit an unknown id")
        return rc

    def H(self, nofixed=True):
        """(colptr, rowidx, val) of the solver's H (N x N) or H_nofixed ((N-3) x (N-3)) as left by the last step()."""
        c = self.counts()
        n = c["N"] - (3 if nofixed else 0)
        nnz = int(self.L.ref_H_nnz(self.h, int(nofixed)))
        colptr = np.zeros(n + 1, np.int32); rowidx = np.zeros(nnz, np.int32); val = np.zeros(nnz, np.float32)
        self.L.ref_get_H(self.h, int(nofixed), _p(colptr), _p(rowidx), _p(val))
        return colptr, rowidx, val

    def b(self, nofixed=True):
        c = self.counts()
        v = np.zeros(c["N"] - (3 if nofixed else 0), np.float32)
        self.L.ref_get_b(self.h, int(nofixed), _p(v))
        return v

    def edge_terms(self, numeric=False):
        c = self.counts()
        eb = np.zeros(c["Eb"], np.float32); jb = np.zeros((c["Eb"], 5), np.float32)
        eo = np.zeros((c["Eo"], 3), np.float32); jo = np.zeros((c["Eo"], 18), np.float32)
        self.L.ref_edge_terms(self.h, int(numeric), _p(eb), _p(jb), _p(eo), _p(jo))
        return eb, jb, eo, jo

    def predict_bearing(self, x, y, th, lx, ly):
        return float(self.L.ref_predict_bearing(self.h, x, y, th, lx, ly))

    def predict_odometry(self, s_xyt, d_xyt):
        s = _f(s_xyt); d = _f(d_xyt); out = np.zeros(3, np.float32)
        self.L.ref_predict_odometry(self.h, _p(s), _p(d), _p(out))
        return out

    def normalized_angle(self, a):
        return float(self.L.ref_normalized_angle(self.h, a))


def smallest_angle(a):
    return float(lib().ref_smallest_angle(a))


def boxplus(xyt, d):
    xyt = _f(xyt); d = _f(d); out = np.zeros(3, np.float32)
    lib().ref_boxplus(_p(xyt), _p(d), _p(out))
    return out
