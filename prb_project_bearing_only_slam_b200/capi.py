"""ctypes binding of include/bos_b200.h -- the same C ABI a cgo/JNI/C++ host would bind.

No compute happens in Python and there is no CPU fallback: if libbos_b200.so is missing the import
fails loudly, and every compute call needs a CUDA device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BOS_LIB_PATH") or os.path.join(_HERE, "libbos_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NCCL, ERR_NOMEM = range(6)
PRECISION_F64, PRECISION_F32 = 0, 1
SOLVER_AUTO, SOLVER_DENSE_CHOLESKY, SOLVER_PCG, SOLVER_SPARSE_CHOLESKY = 0, 1, 2, 3
ROBUST_REFERENCE, ROBUST_IRLS = 0, 1
NCCL_UID_BYTES = 128
IPC_HANDLE_BYTES = 64


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("precision", C.c_int), ("solver", C.c_int), ("dense_max_dim", C.c_int),
                ("kernel_threshold", C.c_double), ("damping", C.c_double), ("pcg_max_iters", C.c_int),
                ("pcg_rtol", C.c_double), ("pcg_variant", C.c_int), ("pcg_precond", C.c_int), ("pcg_coarse_nodes", C.c_int),
                ("pcg_coarse_refresh", C.c_int), ("reserved", C.c_int * 4)]


class Stats(C.Structure):
    _fields_ = [("chi2_bearing", C.c_double), ("chi2_odometry", C.c_double), ("over_bearing", C.c_int64),
                ("over_odometry", C.c_int64), ("delta_inf", C.c_double), ("solver_status", C.c_int),
                ("solver_used", C.c_int), ("pcg_iterations", C.c_int), ("gpu_launches", C.c_int),
                ("ms_linearize", C.c_float), ("ms_solve", C.c_float), ("ms_update", C.c_float), ("ms_allreduce", C.c_float),
                ("precond_used", C.c_int), ("pcg_resolves", C.c_int), ("state_digest", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PatternInfo(C.Structure):
    _fields_ = [("n_hpl", C.c_int64), ("n_hpp_off", C.c_int64), ("csc_n", C.c_int64), ("csc_nnz", C.c_int64),
                ("N", C.c_int64), ("vals_len", C.c_int64)]


# every symbol include/bos_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "bos_default_options", "bos_version", "bos_create", "bos_destroy", "bos_last_error", "bos_set_kernel_threshold", "bos_set_robust_mode",
    "bos_set_damping_factor", "bos_upload_problem", "bos_set_state", "bos_get_state", "bos_linearize", "bos_solve",
    "bos_update", "bos_step", "bos_step_host", "bos_get_stats", "bos_triangulate", "bos_pattern_info_get",
    "bos_download_pattern", "bos_download_blocks", "bos_download_csc", "bos_download_delta", "bos_upload_delta",
    "bos_edge_terms", "bos_host_pattern_create", "bos_host_pattern_destroy", "bos_host_pattern_info",
    "bos_host_pattern_get", "bos_host_pattern_checksum", "bos_host_pattern_skyline", "bos_set_device_setup", "bos_last_setup_ms",
    "bos_pattern_checksum", "bos_device_resolve_ids", "bos_host_edge_shard", "bos_nccl_unique_id", "bos_comm_init", "bos_set_reduce_mode", "bos_peer_export", "bos_peer_open",
    "bos_set_edge_shard", "bos_get_edge_shard", "bos_batch_create", "bos_batch_destroy", "bos_batch_set_states",
    "bos_batch_get_states", "bos_batch_step", "bos_batch_step_device", "bos_batch_last_error",
    "bos_triangulate_landmarks", "bos_eval_bearing_edges", "bos_eval_odometry_edges", "bos_step_lm",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libbos_b200.so is not built (%s); run `python __graft_entry__.py` / "
                              "prb_project_bearing_only_slam_b200/build.py -- there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.bos_default_options.argtypes = [C.POINTER(Options)]
        L.bos_default_options.restype = None
        L.bos_create.argtypes = [C.POINTER(Options), C.POINTER(vp)]
        L.bos_destroy.argtypes = [vp]
        L.bos_last_error.argtypes = [vp]
        L.bos_last_error.restype = C.c_char_p
        L.bos_set_kernel_threshold.argtypes = [vp, dbl]
        L.bos_set_robust_mode.argtypes = [vp, C.c_int]
        L.bos_set_damping_factor.argtypes = [vp, dbl]
        L.bos_upload_problem.argtypes = [vp, i32, i32, i32, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp]
        L.bos_set_state.argtypes = [vp, vp, vp]
        L.bos_get_state.argtypes = [vp, vp, vp]
        for f in ("bos_linearize", "bos_solve", "bos_update"):
            getattr(L, f).argtypes = [vp]
        L.bos_step.argtypes = [vp, C.POINTER(Stats)]
        L.bos_step_host.argtypes = [vp, vp, vp, C.POINTER(Stats)]
        L.bos_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.bos_step_lm.argtypes = [vp, C.POINTER(Stats), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.bos_triangulate.argtypes = [vp, C.POINTER(C.c_int)]
        L.bos_pattern_info_get.argtypes = [vp, C.POINTER(PatternInfo)]
        L.bos_download_pattern.argtypes = [vp] + [vp] * 6
        L.bos_download_blocks.argtypes = [vp] + [vp] * 5
        L.bos_download_csc.argtypes = [vp] + [vp] * 4
        L.bos_download_delta.argtypes = [vp, vp]
        L.bos_upload_delta.argtypes = [vp, vp]
        L.bos_edge_terms.argtypes = [vp] + [vp] * 4
        L.bos_host_pattern_create.argtypes = [i32, i32, i32, i64, vp, vp, i64, vp, vp, C.POINTER(vp)]
        L.bos_host_pattern_destroy.argtypes = [vp]
        L.bos_host_pattern_info.argtypes = [vp, C.POINTER(PatternInfo)]
        L.bos_host_pattern_get.argtypes = [vp] + [vp] * 8
        L.bos_host_pattern_checksum.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.bos_host_pattern_skyline.argtypes = [vp, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.bos_host_edge_shard.argtypes = [i64, i64, i32, i32, vp]
        L.bos_nccl_unique_id.argtypes = [C.c_char_p]
        L.bos_comm_init.argtypes = [vp, i32, i32, C.c_char_p]
        L.bos_set_reduce_mode.argtypes = [vp, i32]
        L.bos_peer_export.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
        L.bos_peer_open.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
        L.bos_set_device_setup.argtypes = [vp, i32]
        L.bos_last_setup_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.bos_pattern_checksum.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.bos_device_resolve_ids.argtypes = [i32, i32, vp, C.c_int64, vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int32)]
        L.bos_set_edge_shard.argtypes = [vp, i32, i32]
        L.bos_get_edge_shard.argtypes = [vp] + [C.POINTER(i64)] * 4
        L.bos_batch_create.argtypes = [C.POINTER(Options), i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, C.POINTER(vp)]
        L.bos_batch_destroy.argtypes = [vp]
        L.bos_batch_set_states.argtypes = [vp, vp, vp]
        L.bos_batch_get_states.argtypes = [vp, vp, vp]
        L.bos_batch_step.argtypes = [vp, vp, vp, vp]
        L.bos_batch_step_device.argtypes = [vp, i32, C.POINTER(C.c_float)]
        L.bos_batch_last_error.argtypes = [vp]
        L.bos_batch_last_error.restype = C.c_char_p
        L.bos_triangulate_landmarks.argtypes = [C.POINTER(Options), i32, vp, i64, vp, vp, vp, i32, vp, C.POINTER(C.c_int)]
        L.bos_eval_bearing_edges.argtypes = [C.POINTER(Options), i64, vp, vp, vp, vp, vp]
        L.bos_eval_odometry_edges.argtypes = [C.POINTER(Options), i64, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


class BosError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "bos error %d: %s" % (code, msg))
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def default_options(**kw):
    o = Options()
    lib().bos_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Context:
    """Thin object wrapper over a bos_ctx*; one per GPU, one host thread at a time."""

    def __init__(self, **opts):
        self.L = lib()
        self.opts = default_options(**opts)
        self.h = C.c_void_p()
        rc = self.L.bos_create(C.byref(self.opts), C.byref(self.h))
        if rc != OK:
            raise BosError(rc, "bos_create failed (no CUDA device? there is no CPU fallback)")
        self.NP = self.NL = self.Eb = self.Eo = 0

    def close(self):
        if self.h:
            self.L.bos_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise BosError(rc, (self.L.bos_last_error(self.h) or b"").decode())

    def set_kernel_threshold(self, kt):
        self._ck(self.L.bos_set_kernel_threshold(self.h, float(kt)))

    def set_robust_mode(self, mode):
        """0 = the reference's robust kernel (error scaled), 1 = IRLS (Omega scaled)."""
        self._ck(self.L.bos_set_robust_mode(self.h, int(mode)))

    def set_damping_factor(self, df):
        self._ck(self.L.bos_set_damping_factor(self.h, float(df)))

    def upload_problem(self, NP, NL, fixed_stix, b_pose, b_lm, b_z, b_omega, o_src, o_dst, o_z, o_omega):
        b_pose = _c(b_pose, np.int32); b_lm = _c(b_lm, np.int32); b_z = _c(b_z, np.float64); b_omega = _c(b_omega, np.float64)
        o_src = _c(o_src, np.int32); o_dst = _c(o_dst, np.int32); o_z = _c(o_z, np.float64); o_omega = _c(o_omega, np.float64)
        self.NP, self.NL, self.Eb, self.Eo = int(NP), int(NL), len(b_z), len(o_src)
        self._ck(self.L.bos_upload_problem(self.h, int(NP), int(NL), int(fixed_stix), self.Eb, _ptr(b_pose), _ptr(b_lm), _ptr(b_z),
                                           _ptr(b_omega), self.Eo, _ptr(o_src), _ptr(o_dst), _ptr(o_z), _ptr(o_omega)))

    def set_state(self, poses_xycs=None, lms_xy=None):
        p = _c(poses_xycs, np.float64); l = _c(lms_xy, np.float64)
        self._ck(self.L.bos_set_state(self.h, _ptr(p), _ptr(l)))

    def get_state(self):
        p = np.zeros((self.NP, 4)); l = np.zeros((self.NL, 2))
        self._ck(self.L.bos_get_state(self.h, _ptr(p), _ptr(l)))
        return p, l

    def linearize(self):
        self._ck(self.L.bos_linearize(self.h))

    def solve(self):
        self._ck(self.L.bos_solve(self.h))

    def update(self):
        self._ck(self.L.bos_update(self.h))

    def step(self):
        s = Stats()
        self._ck(self.L.bos_step(self.h, C.byref(s)))
        return s

    def step_lm(self):
        """Opt-in Levenberg-Marquardt iteration (extension): returns (stats of the GN step taken, chi2 after, accepted, next damping)."""
        s = Stats(); after = C.c_double(0); ok = C.c_int(0); damp = C.c_double(0)
        self._ck(self.L.bos_step_lm(self.h, C.byref(s), C.byref(after), C.byref(ok), C.byref(damp)))
        return s, after.value, bool(ok.value), damp.value

    def step_host(self, poses_xycs, lms_xy):
        """poses/lms must be C-contiguous float64 arrays; updated in place."""
        s = Stats()
        self._ck(self.L.bos_step_host(self.h, _ptr(poses_xycs), _ptr(lms_xy), C.byref(s)))
        return s

    def stats(self):
        s = Stats()
        self._ck(self.L.bos_get_stats(self.h, C.byref(s)))
        return s

    def triangulate(self):
        n = C.c_int(0)
        self._ck(self.L.bos_triangulate(self.h, C.byref(n)))
        return n.value

    def pattern_info(self):
        pi = PatternInfo()
        self._ck(self.L.bos_pattern_info_get(self.h, C.byref(pi)))
        return pi

    def pattern(self):
        pi = self.pattern_info()
        hp = np.zeros(pi.n_hpl, np.int32); hl = np.zeros(pi.n_hpl, np.int32)
        lo = np.zeros(pi.n_hpp_off, np.int32); hi = np.zeros(pi.n_hpp_off, np.int32)
        bs = np.zeros(self.Eb, np.int64); os_ = np.zeros(self.Eo, np.int64)
        self._ck(self.L.bos_download_pattern(self.h, _ptr(hp), _ptr(hl), _ptr(lo), _ptr(hi), _ptr(bs), _ptr(os_)))
        return dict(hpl_pose=hp, hpl_lm=hl, off_lo=lo, off_hi=hi, b_slot=bs, o_slot=os_)

    def blocks(self):
        pi = self.pattern_info()
        Hpp = np.zeros((self.NP, 9)); Hll = np.zeros((self.NL, 4)); Hpl = np.zeros((pi.n_hpl, 6))
        Hoff = np.zeros((pi.n_hpp_off, 9)); b = np.zeros(pi.N)
        self._ck(self.L.bos_download_blocks(self.h, _ptr(Hpp), _ptr(Hll), _ptr(Hpl), _ptr(Hoff), _ptr(b)))
        return dict(Hpp=Hpp, Hll=Hll, Hpl=Hpl, Hoff=Hoff, b=b)

    def csc(self, values=True):
        pi = self.pattern_info()
        colptr = np.zeros(pi.csc_n + 1, np.int32); rowidx = np.zeros(pi.csc_nnz, np.int32)
        val = np.zeros(pi.csc_nnz) if values else None
        b = np.zeros(pi.csc_n) if values else None
        self._ck(self.L.bos_download_csc(self.h, _ptr(colptr), _ptr(rowidx), _ptr(val), _ptr(b)))
        return colptr, rowidx, val, b

    def delta(self):
        d = np.zeros(self.pattern_info().N)
        self._ck(self.L.bos_download_delta(self.h, _ptr(d)))
        return d

    def upload_delta(self, d):
        d = _c(d, np.float64)
        self._ck(self.L.bos_upload_delta(self.h, _ptr(d)))

    def edge_terms(self):
        eb = np.zeros(self.Eb); jb = np.zeros((self.Eb, 5)); eo = np.zeros((self.Eo, 3)); jo = np.zeros((self.Eo, 18))
        self._ck(self.L.bos_edge_terms(self.h, _ptr(eb), _ptr(jb), _ptr(eo), _ptr(jo)))
        return eb, jb, eo, jo

    def comm_init(self, rank, nranks, uid):
        self._ck(self.L.bos_comm_init(self.h, int(rank), int(nranks), uid))

    def set_reduce_mode(self, mode):
        self._ck(self.L.bos_set_reduce_mode(self.h, int(mode)))

    def peer_export(self):
        """(CUDA IPC handle of the value buffer, byte offset inside its allocation) -- reduce_mode 4, see bos_peer_export."""
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        off = C.c_int64()
        self._ck(self.L.bos_peer_export(self.h, buf, C.byref(off)))
        return buf.raw, int(off.value)

    def peer_open(self, handles, offsets):
        """handles / offsets of ALL ranks, indexed by rank (what every rank's peer_export returned)."""
        blob = b"".join(handles)
        arr = (C.c_int64 * len(offsets))(*[int(o) for o in offsets])
        self._ck(self.L.bos_peer_open(self.h, blob, arr))

    def peer_connect(self, dist):
        """Exchange the IPC handles over torch.distributed and open the peers' value buffers (collective)."""
        mine = self.peer_export()
        everyone = [None] * dist.get_world_size()
        dist.all_gather_object(everyone, mine)
        self.peer_open([h for h, _ in everyone], [o for _, o in everyone])

    def set_device_setup(self, on=True):
        """Build the bearing-edge core of the pattern on the device at the next upload_problem: True / False, None = the
        library's default (on the device from 200 000 bearing edges on)."""
        self._ck(self.L.bos_set_device_setup(self.h, -1 if on is None else (1 if on else 0)))

    def last_setup_ms(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.bos_last_setup_ms(self.h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def pattern_checksum(self):
        out = C.c_uint64()
        self._ck(self.L.bos_pattern_checksum(self.h, C.byref(out)))
        return int(out.value)

    def set_edge_shard(self, rank, nranks):
        self._ck(self.L.bos_set_edge_shard(self.h, int(rank), int(nranks)))

    def edge_shard(self):
        v = [C.c_int64() for _ in range(4)]
        self._ck(self.L.bos_get_edge_shard(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)


def nccl_unique_id():
    buf = C.create_string_buffer(NCCL_UID_BYTES)
    rc = lib().bos_nccl_unique_id(buf)
    if rc != OK:
        raise BosError(rc, "bos_nccl_unique_id failed")
    return buf.raw


def device_resolve_ids(pose_ids, b_pose_id, b_lm_id, o_src_id, o_dst_id, device=0):
    """id -> stix of every edge end point on the device; returns (b_pose, b_lm, o_src, o_dst, lm_ids).  Unknown pose id: BosError."""
    L = lib()
    pose_ids = _c(pose_ids, np.int32); b_pose_id = _c(b_pose_id, np.int32); b_lm_id = _c(b_lm_id, np.int32)
    o_src_id = _c(o_src_id, np.int32); o_dst_id = _c(o_dst_id, np.int32)
    Eb, Eo = len(b_pose_id), len(o_src_id)
    bp = np.zeros(Eb, np.int32); bl = np.zeros(Eb, np.int32); os_ = np.zeros(Eo, np.int32); od = np.zeros(Eo, np.int32)
    lm = np.zeros(max(Eb, 1), np.int32)
    nl = C.c_int32(0)
    rc = L.bos_device_resolve_ids(int(device), len(pose_ids), _ptr(pose_ids), Eb, _ptr(b_pose_id), _ptr(b_lm_id), Eo, _ptr(o_src_id), _ptr(o_dst_id),
                                  _ptr(bp), _ptr(bl), _ptr(os_), _ptr(od), _ptr(lm), C.byref(nl))
    if rc != OK:
        raise BosError(rc, "bos_device_resolve_ids")
    return bp, bl, os_, od, lm[:nl.value].copy()


class HostPattern:
    """Host-only pattern builder view (no device needed)."""

    def __init__(self, NP, NL, fixed_stix, b_pose, b_lm, o_src, o_dst):
        self.L = lib()
        b_pose = _c(b_pose, np.int32); b_lm = _c(b_lm, np.int32); o_src = _c(o_src, np.int32); o_dst = _c(o_dst, np.int32)
        self.Eb, self.Eo = len(b_pose), len(o_src)
        self.h = C.c_void_p()
        rc = self.L.bos_host_pattern_create(int(NP), int(NL), int(fixed_stix), self.Eb, _ptr(b_pose), _ptr(b_lm), self.Eo,
                                            _ptr(o_src), _ptr(o_dst), C.byref(self.h))
        if rc != OK:
            raise BosError(rc, "bos_host_pattern_create rejected the problem")

    def __del__(self):
        try:
            if self.h:
                self.L.bos_host_pattern_destroy(self.h)
        except Exception:
            pass

    def info(self):
        pi = PatternInfo()
        self.L.bos_host_pattern_info(self.h, C.byref(pi))
        return pi

    def skyline(self):
        """Symbolic phase of the skyline Cholesky: (panel_end[int32], rows per column, stored fraction of the lower triangle)."""
        npan, W, fill = C.c_int32(), C.c_int32(), C.c_double()
        self.L.bos_host_pattern_skyline(self.h, None, C.byref(npan), C.byref(W), C.byref(fill))
        pe = np.zeros(npan.value, np.int32)
        rc = self.L.bos_host_pattern_skyline(self.h, _ptr(pe), C.byref(npan), C.byref(W), C.byref(fill))
        if rc != OK:
            raise BosError(rc, "bos_host_pattern_skyline")
        return pe, int(W.value), float(fill.value)

    def checksum(self):
        out = C.c_uint64()
        rc = self.L.bos_host_pattern_checksum(self.h, C.byref(out))
        if rc != OK:
            raise BosError(rc, "bos_host_pattern_checksum")
        return int(out.value)

    def get(self):
        pi = self.info()
        hp = np.zeros(pi.n_hpl, np.int32); hl = np.zeros(pi.n_hpl, np.int32)
        lo = np.zeros(pi.n_hpp_off, np.int32); hi = np.zeros(pi.n_hpp_off, np.int32)
        bs = np.zeros(self.Eb, np.int64); os_ = np.zeros(self.Eo, np.int64)
        colptr = np.zeros(pi.csc_n + 1, np.int32); rowidx = np.zeros(pi.csc_nnz, np.int32)
        rc = self.L.bos_host_pattern_get(self.h, _ptr(hp), _ptr(hl), _ptr(lo), _ptr(hi), _ptr(bs), _ptr(os_), _ptr(colptr), _ptr(rowidx))
        if rc != OK:
            raise BosError(rc, "bos_host_pattern_get")
        return dict(hpl_pose=hp, hpl_lm=hl, off_lo=lo, off_hi=hi, b_slot=bs, o_slot=os_, csc_colptr=colptr, csc_rowidx=rowidx)


def host_edge_shard(Eb, Eo, rank, nranks):
    out = np.zeros(4, np.int64)
    rc = lib().bos_host_edge_shard(int(Eb), int(Eo), int(rank), int(nranks), _ptr(out))
    if rc != OK:
        raise BosError(rc, "bos_host_edge_shard")
    return tuple(int(x) for x in out)


class Batch:
    def __init__(self, nprob, NP, NL, fixed_stix, b_pose, b_lm, b_z, b_omega, o_src, o_dst, o_z, o_omega, **opts):
        self.L = lib()
        self.opts = default_options(**opts)
        b_pose = _c(b_pose, np.int32); b_lm = _c(b_lm, np.int32); b_z = _c(b_z, np.float64); b_omega = _c(b_omega, np.float64)
        o_src = _c(o_src, np.int32); o_dst = _c(o_dst, np.int32); o_z = _c(o_z, np.float64); o_omega = _c(o_omega, np.float64)
        self.nprob, self.NP, self.NL = int(nprob), int(NP), int(NL)
        self.h = C.c_void_p()
        rc = self.L.bos_batch_create(C.byref(self.opts), self.nprob, self.NP, self.NL, int(fixed_stix), len(b_pose), _ptr(b_pose),
                                     _ptr(b_lm), _ptr(b_z), _ptr(b_omega), len(o_src), _ptr(o_src), _ptr(o_dst), _ptr(o_z),
                                     _ptr(o_omega), C.byref(self.h))
        if rc != OK:
            raise BosError(rc, "bos_batch_create failed")

    def close(self):
        if self.h:
            self.L.bos_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise BosError(rc, (self.L.bos_batch_last_error(self.h) or b"").decode())

    def set_states(self, poses, lms):
        p = _c(poses, np.float64); l = _c(lms, np.float64)
        self._ck(self.L.bos_batch_set_states(self.h, _ptr(p), _ptr(l)))

    def get_states(self):
        p = np.zeros((self.nprob, self.NP, 4)); l = np.zeros((self.nprob, self.NL, 2))
        self._ck(self.L.bos_batch_get_states(self.h, _ptr(p), _ptr(l)))
        return p, l

    def step(self):
        chi = np.zeros((self.nprob, 2)); d = np.zeros(self.nprob); st = np.zeros(self.nprob, np.int32)
        self._ck(self.L.bos_batch_step(self.h, _ptr(chi), _ptr(d), _ptr(st)))
        return chi, d, st

    def step_device(self, n_steps):
        ms = C.c_float(0)
        self._ck(self.L.bos_batch_step_device(self.h, int(n_steps), C.byref(ms)))
        return ms.value


def synth_world(n_poses, n_landmarks, target_bearing_edges, seed=0xB0500000, **kw):
    """The synthetic-world generator lives in its own library (synth/, not part of libbos_b200.so); kept here as a forwarder."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from synth import synth_world as gen
    return gen(n_poses, n_landmarks, target_bearing_edges, seed=seed, **kw)
