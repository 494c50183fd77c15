// ctx.cu -- the C ABI of include/bos_b200.h: context, problem upload, phase sequencing, NCCL plumbing.
// No CPU fallback lives here: every compute entry point launches CUDA kernels or fails.
#include "../../include/bos_b200.h"
#include "bos_internal.h"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace bos;

namespace {

struct DevAlloc {
    std::vector<void*> ptrs;
    ~DevAlloc() { release(); }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
    template <typename T>
    T* get(size_t count) {
        void* p = nullptr;
        if (count == 0) count = 1;
        if (cudaMalloc(&p, count * sizeof(T)) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
    template <typename T>
    T* upload(const std::vector<T>& v) {
        T* p = get<T>(v.size());
        if (p && !v.empty() && cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        return p;
    }
};

// ---- NCCL through dlopen: the single-GPU path never depends on it -------------------------------------
typedef struct { char internal[BOS_NCCL_UID_BYTES]; } nccl_uid_t;
typedef void* nccl_comm_t;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.GetUniqueId = (int (*)(nccl_uid_t*))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_uid_t, int))dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (int (*)(nccl_comm_t))dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
            api.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))dlsym(api.handle, "ncclAllGather");
            api.Broadcast = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(api.handle, "ncclBroadcast");
            api.GroupStart = (int (*)())dlsym(api.handle, "ncclGroupStart");
            api.GroupEnd = (int (*)())dlsym(api.handle, "ncclGroupEnd");
            api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.Broadcast;
        }
    }
    return api;
}
constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

}  // namespace

struct bos_ctx {
    bos_options opt;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;
    bool have_problem = false, linearized = false, solved = false;
    bool stepping_host = false;          // bos_step_host: the uploaded state is the one the previous step produced
    bool delta_valid = false;            // set by bos_solve / bos_upload_delta, consumed by bos_update
    HostPattern P;
    DevAlloc mem;
    Dev<double> dd;
    Dev<float> df;
    DenseWork<double> dwd; DenseWork<float> dwf;
    DenseWork<double> swd; DenseWork<float> swf;      // skyline flavour (BOS_SOLVER_SPARSE_CHOLESKY)
    bool sky_ready = false;
    PcgWork<double> pwd; PcgWork<float> pwf;
    bool dense_ready = false, pcg_ready = false;
    void* edge_scratch = nullptr;
    int* d_single_obs = nullptr;
    size_t vals_len = 0, vals_prefix = 0, hpl_padded = 0;
    bos_stats stats;
    int solver_used = 0;
    // multi-GPU
    int rank = 0, nranks = 1, reduce_mode = 0;
    int robust_mode = 0;          // 0 = reference robust kernel, 1 = IRLS (bos_set_robust_mode)
    int device_setup = -1;        // 1: the bearing-edge core of the pattern is built on the device, 0: on host threads, -1: device from 200 k edges on (bos_set_device_setup)
    bool device_setup_used = false;
    double* pinned = nullptr;     // 64 doubles of pinned host memory: statistics and solver scalars are read back through it
    double setup_ms[2] = {0.0, 0.0};   // last upload: device core, host remainder
    nccl_comm_t comm = nullptr;
    // reduce_mode 4: the ranks' value buffers mapped into this process through CUDA IPC (bos_peer_export / bos_peer_open)
    size_t tail_off = 0;                          // byte offset of the statistics / barrier tail inside the value-buffer allocation
    unsigned char* peer_base[kMaxPeers] = {};     // value buffer of every rank (own included) in this process' address space
    void* peer_opened[kMaxPeers] = {};            // what cudaIpcOpenMemHandle returned (to close)
    bool peers_open = false;
    unsigned long long peer_epoch = 0;
    void* peer_scratch = nullptr;                 // reduce_mode 5: totals pulled from the peers, committed after the second barrier
    ShardRange shard;
    std::vector<int> own_p0;              // [nranks + 1] first pose owned by each rank's tiles (a pose belongs to the tile its run starts in)
    int shard_chunk_b = 0;
    int launches = 0;
    bool pcg_bad = false, pcg_capped = false;
    std::vector<double> b_omega_sorted;   // bearing omegas in sorted-edge order (host copy for the PCG setup)
    void* lm_pose_bak = nullptr;          // state backup of bos_step_lm
    void* lm_lm_bak = nullptr;

    bool f64() const { return opt.precision == BOS_PRECISION_F64; }
};

namespace {

int fail(bos_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
#define CUDA_OK(c, call)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) return fail(c, BOS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

template <typename S> Dev<S>& dev(bos_ctx* c);
template <> Dev<double>& dev<double>(bos_ctx* c) { return c->dd; }
template <> Dev<float>& dev<float>(bos_ctx* c) { return c->df; }
template <typename S> DenseWork<S>& dwork(bos_ctx* c);
template <> DenseWork<double>& dwork<double>(bos_ctx* c) { return c->dwd; }
template <> DenseWork<float>& dwork<float>(bos_ctx* c) { return c->dwf; }
template <typename S> DenseWork<S>& swork(bos_ctx* c);
template <> DenseWork<double>& swork<double>(bos_ctx* c) { return c->swd; }
template <> DenseWork<float>& swork<float>(bos_ctx* c) { return c->swf; }
template <typename S> PcgWork<S>& pwork(bos_ctx* c);
template <> PcgWork<double>& pwork<double>(bos_ctx* c) { return c->pwd; }
template <> PcgWork<float>& pwork<float>(bos_ctx* c) { return c->pwf; }

template <typename S>
std::vector<S> narrow(const double* src, size_t n) {
    std::vector<S> v(n);
    for (size_t i = 0; i < n; i++) v[i] = (S)src[i];
    return v;
}

void shard_ranges(int64_t Eb, int64_t Eo, int r, int R, int64_t out[4], int64_t* chunk_b) {
    const int64_t cb = (((Eb + R - 1) / R) + kLinTile - 1) / kLinTile * kLinTile, co = (Eo + R - 1) / R;  // bearing shards are whole tiles
    if (chunk_b) *chunk_b = cb;
    out[0] = std::min(Eb, r * cb); out[1] = std::min(Eb, (r + 1) * cb);
    out[2] = std::min(Eo, r * co); out[3] = std::min(Eo, (r + 1) * co);
}
void compute_shard(bos_ctx* c) {
    int64_t o[4], cb;
    shard_ranges(c->P.Eb, c->P.Eo, c->rank, c->nranks, o, &cb);
    c->shard_chunk_b = (int)cb;
    c->shard.b_begin = (int)o[0]; c->shard.b_end = (int)o[1];
    c->shard.o_begin = (int)o[2]; c->shard.o_end = (int)o[3];
    c->own_p0.assign(c->nranks + 1, c->P.NP);
    if ((int)c->P.epose_ptr.size() < c->P.NP + 1) return;      // no problem uploaded yet
    for (int r = 0; r < c->nranks; r++) {
        int64_t q[4];
        shard_ranges(c->P.Eb, c->P.Eo, r, c->nranks, q, nullptr);
        // first pose whose run starts at or after the rank's first edge; rank 0 starts at pose 0, trailing edge-free poses go to the last
        // rank that has edges (the tile with tb == Eb), ranks without edges own nothing
        c->own_p0[r] = (q[0] >= c->P.Eb && r > 0) ? c->P.NP
                       : (int)(std::lower_bound(c->P.epose_ptr.begin(), c->P.epose_ptr.begin() + c->P.NP + 1, (int)q[0]) - c->P.epose_ptr.begin());
    }
    c->own_p0[0] = 0;
}

template <typename S>
int upload_impl(bos_ctx* c, const double* b_z, const double* b_omega, const double* o_z, const double* o_omega) {
    HostPattern& P = c->P;
    Dev<S>& d = dev<S>(c);
    d = Dev<S>();
    DevAlloc& m = c->mem;
    d.NP = P.NP; d.NL = P.NL; d.fixed = P.fixed; d.Eb = P.Eb; d.Eo = P.Eo; d.N = P.N;
    d.n_hpl = (int)P.slot_pose.size(); d.n_off = (int)P.off_lo.size();
    // bearing SoA in sorted order, padded to a multiple of 4 edges (the kernel reads 4 per thread); padding has omega = 0
    d.Eb_pad = (P.Eb + kLinTile - 1) / kLinTile * kLinTile + kLinTile;   // whole tiles: the bearing kernel fetches tiles by TMA bulk copies
    d.hpl_ld = (d.n_hpl + 3) / 4 * 4 + 8 * kLinTile;   // room for up to 8 tile-padded rank shards   // room for the padded in-place allgather of rank shards
    std::vector<S> bz(d.Eb_pad, S(0)), bom(d.Eb_pad, S(0));
    std::vector<int> bpose(d.Eb_pad, P.Eb ? P.b_pose[P.Eb - 1] : 0), blm(d.Eb_pad, 0), bslot(d.Eb_pad, 0);
    c->b_omega_sorted.assign(P.Eb, 1.0);
    for (int k = 0; k < P.Eb; k++) {
        bz[k] = (S)b_z[P.b_perm[k]];
        bom[k] = b_omega ? (S)b_omega[P.b_perm[k]] : S(1);
        if (b_omega) c->b_omega_sorted[k] = (double)bom[k];
        bpose[k] = P.b_pose[k]; blm[k] = P.b_lm[k]; bslot[k] = P.b_slot[k];
    }
    d.has_shared_off = P.has_shared_off ? 1 : 0;
    std::vector<S> oz((size_t)3 * P.Eo), oom((size_t)6 * P.Eo);
    static const int up[6] = {0, 1, 2, 4, 5, 8};
    for (int e = 0; e < P.Eo; e++) {
        for (int k = 0; k < 3; k++) oz[(size_t)k * P.Eo + e] = (S)o_z[3 * (size_t)e + k];
        for (int k = 0; k < 6; k++) oom[(size_t)k * P.Eo + e] = (S)o_omega[9 * (size_t)e + up[k]];
    }
#define UP(field, vec)                                   \
    d.field = m.upload(vec);                             \
    if (!d.field) return fail(c, BOS_ERR_NOMEM, "device allocation failed: " #field);
    UP(b_pose, bpose) UP(b_lm, blm) UP(b_z, bz) UP(b_om, bom) UP(b_perm, P.b_perm)
    if (!P.slots_identity) { UP(b_slot, bslot) }
    UP(o_src, P.o_src) UP(o_dst, P.o_dst) UP(o_z, oz) UP(o_om, oom) UP(o_slot, P.o_slot)
    UP(oe_ptr, P.oe_ptr) UP(oe_edge, P.oe_edge) UP(oe_other, P.oe_other) UP(o_shared, P.o_shared)
    {
        std::vector<int> oe2(2 * ((size_t)P.NP + 8), -1);   // padded: the bearing kernel fetches it in whole groups of four poses
        for (int i = 0; i < P.NP; i++)
            for (int k = 0; k < 2 && P.oe_ptr[i] + k < P.oe_ptr[i + 1]; k++) oe2[2 * (size_t)i + k] = P.oe_edge[P.oe_ptr[i] + k];
        UP(oe2, oe2)
    }
    UP(slot_pose, P.slot_pose) UP(slot_lm, P.slot_lm) UP(pose_ptr, P.pose_ptr) UP(lm_ptr, P.lm_ptr)
    UP(lm_order, P.lm_order) UP(lm_order_pose, P.lm_order_pose) UP(lm_order_lm, P.lm_order_lm)
    UP(pp_ptr, P.pp_ptr) UP(pp_nbr, P.pp_nbr) UP(pp_slot, P.pp_slot) UP(off_lo, P.off_lo) UP(off_hi, P.off_hi)
    UP(tri_ptr, P.tri_ptr) UP(tri_edge, P.tri_edge)
    {
        std::vector<int> ep(P.epose_ptr);
        ep.resize(ep.size() + 16, P.Eb);   // bulk copies read whole 16-byte groups
        UP(epose_ptr, ep)
    }
    UP(tile_ptr, P.tile_ptr) UP(tg_lm, P.tg_lm) UP(tg_eptr, P.tg_eptr)
    {
        std::vector<unsigned short> tge(d.Eb_pad, 0);
        std::copy(P.tg_edge.begin(), P.tg_edge.end(), tge.begin());
        UP(tg_edge, tge)
        // per-tile metadata in TMA-friendly form: group lists padded to multiples of 8 entries, one int4 header per tile
        const int ntiles = (int)P.tile_ptr.size() - 1;
        std::vector<int> meta(4 * (size_t)std::max(ntiles, 1), 0), glm;
        std::vector<unsigned short> gep;
        for (int t = 0; t < ntiles; t++) {
            const int g0 = P.tile_ptr[t], g1 = P.tile_ptr[t + 1], ng = g1 - g0, ta = t * kLinTile, tb = std::min(P.Eb, ta + kLinTile);
            // pose range the tile walks: the runs it intersects plus the poses it OWNS (run starts in the tile; edge-free poses included)
            const int own0 = (int)(std::lower_bound(P.epose_ptr.begin(), P.epose_ptr.begin() + P.NP + 1, ta) - P.epose_ptr.begin());
            const int own1 = (t + 1 < ntiles) ? (int)(std::lower_bound(P.epose_ptr.begin(), P.epose_ptr.begin() + P.NP + 1, ta + kLinTile) - P.epose_ptr.begin()) : P.NP;
            meta[4 * t] = (int)glm.size(); meta[4 * t + 1] = ng;
            meta[4 * t + 2] = std::min(P.b_pose[ta], own0); meta[4 * t + 3] = std::max(P.b_pose[tb - 1], own1 - 1);
            const int cnt = (ng + 1 + 7) & ~7;
            for (int j = 0; j < cnt; j++) {
                glm.push_back(j < ng ? P.tg_lm[g0 + j] : 0);
                gep.push_back((unsigned short)(j <= ng ? P.tg_eptr[g0 + j] - ta : 0));
            }
        }
        glm.resize(glm.size() + 8, 0); gep.resize(gep.size() + 8, 0);
        UP(tile_meta, meta) UP(tgp_lm, glm) UP(tgp_eptr, gep)
        std::vector<int> cut;
        for (int p = 0; p < P.NP; p++) {
            const int ra = P.epose_ptr[p], rb = P.epose_ptr[p + 1];
            if (rb > ra && ra / kLinTile != (rb - 1) / kLinTile) cut.push_back(p);
        }
        d.n_cut = (int)cut.size();
        cut.resize(cut.size() + 1, 0);
        UP(cut_pose, cut)
    }
    UP(pl_lm_id, P.pl_lm_id) UP(ell_Loff, P.ell_Loff) UP(ell_Lpose, P.ell_Lpose)
    UP(pc_row_pose, P.pc_row_pose) UP(pc_goff, P.pc_goff) UP(pc_cl_ptr, P.pc_cl_ptr) UP(pc_cl_row, P.pc_cl_row) UP(pc_loc, P.pc_loc)
    UP(pc_emap, P.pc_emap) UP(pc_nbr, P.pc_nbr) UP(pc_nslot, P.pc_nslot) UP(pc_ncnt, P.pc_ncnt)
    if (P.pc_ok) {
        std::vector<int> a = P.lc_gptr, b = P.lc_goff, e = P.sh_ptr, f = P.sh_src, se = P.sh_ell;
        se.resize(se.size() + 1, -1);
        UP(sh_ell, se)
        std::vector<unsigned short> r = P.lc_row, lk = P.lc_k;
        lk.resize(lk.size() + 32, 0xffff);
        UP(lc_k, lk)
        std::vector<unsigned char> g = P.sh_first;
        a.resize(a.size() + 1, 0); b.resize(b.size() + 1, 0); e.resize(e.size() + 1, 0); f.resize(f.size() + 1, 0); r.resize(r.size() + 32, 0xffff); g.resize(g.size() + 1, 0);
        UP(lc_gptr, a) UP(lc_goff, b) UP(lc_row, r) UP(sh_ptr, e) UP(sh_src, f) UP(sh_first, g)
        d.n_q = (int)P.pc_cl_row.size();
    }
    d.n_clm = (int)P.pl_lm_id.size(); d.nLg = (int)P.ell_Loff.size() - 1;
    d.nLs = (long long)P.ell_Lmap.size(); d.nPs = (long long)P.pc_loc.size();
    d.pc_chunks = P.pc_chunks; d.pc_cp = P.pc_cp; d.pc_ok = P.pc_ok ? 1 : 0;
    d.pc_cl_max = 0; d.pc_slots_max = 0;
    for (int q = 0; q < P.pc_chunks; q++) {
        d.pc_cl_max = std::max(d.pc_cl_max, P.pc_cl_ptr[q + 1] - P.pc_cl_ptr[q]);
        const int gpc = P.pc_cp / 32;
        d.pc_slots_max = std::max(d.pc_slots_max, (P.pc_goff[(size_t)(q + 1) * gpc] - P.pc_goff[(size_t)q * gpc]) * 32);
    }
#undef UP
    d.pose = m.get<S>(4 * (size_t)P.NP);
    d.lm = m.get<S>(2 * (size_t)std::max(P.NL, 1));
    d.theta = m.get<S>((size_t)P.NP);
    c->vals_prefix = ((size_t)P.N + 6 * (size_t)P.NP + 3 * (size_t)P.NL + 9 * (size_t)d.n_off + 7) / 8 * 8;  // Hpl planes 32-byte aligned
    c->hpl_padded = 6 * (size_t)d.hpl_ld;
    c->vals_len = c->vals_prefix + c->hpl_padded;
    // one allocation: [ values | tail: statistics, odometry-share scratch, peer barrier slots, error flag ] -- the window the other ranks
    // map through CUDA IPC in reduce_mode 4
    c->tail_off = ((c->vals_prefix + c->hpl_padded) * sizeof(S) + 15) / 16 * 16;
    unsigned char* vbase = m.get<unsigned char>(c->tail_off + kTailWords * 8);
    d.vals = reinterpret_cast<S*>(vbase);
    d.stats = vbase ? reinterpret_cast<double*>(vbase + c->tail_off) : nullptr;
    d.stats_k2 = d.stats;
    d.npeer = 0;
    d.Mv = m.get<S>(9 * (size_t)std::max(P.Eo, 1));
    d.delta = m.get<S>((size_t)P.N);
    c->d_single_obs = m.get<int>(1);
    if (!d.pose || !d.lm || !d.theta || !d.vals || !d.stats || !d.Mv || !d.delta || !c->d_single_obs) return fail(c, BOS_ERR_NOMEM, "device allocation failed");
    d.b = d.vals;
    d.Hpp = d.b + P.N;
    d.Hll = d.Hpp + 6 * (size_t)P.NP;
    d.Hoff = d.Hll + 3 * (size_t)P.NL;
    d.Hpl = d.vals + c->vals_prefix;
    // on the context's own (non-blocking) stream: a legacy-stream memset is not ordered with later copies on that stream
    CUDA_OK(c, cudaMemsetAsync(d.vals, 0, (c->vals_prefix + c->hpl_padded) * sizeof(S), c->stream));
    CUDA_OK(c, cudaMemsetAsync(d.delta, 0, (size_t)P.N * sizeof(S), c->stream));
    CUDA_OK(c, cudaMemsetAsync(d.stats, 0, kTailWords * 8, c->stream));
    CUDA_OK(c, cudaMemsetAsync(d.pose, 0, 4 * (size_t)P.NP * sizeof(S), c->stream));
    CUDA_OK(c, cudaMemsetAsync(d.theta, 0, (size_t)P.NP * sizeof(S), c->stream));
    CUDA_OK(c, cudaMemsetAsync(d.lm, 0, 2 * (size_t)std::max(P.NL, 1) * sizeof(S), c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BOS_OK;
}

template <typename S>
int ensure_dense(bos_ctx* c) {
    if (c->dense_ready) return BOS_OK;
    Dev<S>& d = dev<S>(c);
    DenseWork<S>& w = dwork<S>(c);
    w.n = 3 * d.NP;
    w.Smat = c->mem.get<S>((size_t)w.n * w.n);
    w.g = c->mem.get<S>((size_t)w.n);
    w.hllinv = c->mem.get<S>(3 * (size_t)std::max(d.NL, 1));
    w.ul = c->mem.get<S>(2 * (size_t)std::max(d.NL, 1));
    w.tl = c->mem.get<S>(2 * (size_t)std::max(d.NL, 1));
    w.tl_blk = c->mem.get<S>(64);
    if (!w.Smat || !w.g || !w.hllinv || !w.ul || !w.tl || !w.tl_blk) return fail(c, BOS_ERR_NOMEM, "dense workspace allocation failed");
    c->dense_ready = true;
    return BOS_OK;
}
// Symbolic phase of the skyline Cholesky (the analogue of SimplicialLDLT::analyzePattern, slam/solver.cpp:77-80), once per uploaded pattern: pose p is coupled with every pose that shares a landmark or an odometry
// edge with it; in pose order the factor's fill stays inside the monotone envelope, so column j needs rows up to the largest coupled row of
// any column <= j.  Row limits are kept per 64-column panel; W = the tallest window any outer panel of the factorisation touches.
template <typename S>
void skyline_symbolic(const HostPattern& P, DenseWork<S>& w) {
    const int NP = P.NP, n = 3 * NP;
    std::vector<int> lastpose(NP);
    for (int p = 0; p < NP; p++) lastpose[p] = p;
    for (int l = 0; l < P.NL; l++) {
        const int a = P.lm_ptr[l], b = P.lm_ptr[l + 1];
        if (b <= a) continue;
        const int pmax = P.lm_order_pose[b - 1];      // ascending pose inside a landmark
        for (int k = a; k < b; k++) lastpose[P.lm_order_pose[k]] = std::max(lastpose[P.lm_order_pose[k]], pmax);
    }
    for (size_t k = 0; k < P.off_lo.size(); k++) lastpose[P.off_lo[k]] = std::max(lastpose[P.off_lo[k]], P.off_hi[k]);
    const int npanels = (n + kDenseNB - 1) / kDenseNB;
    w.sky_panel_end.assign(npanels, 0);
    int run = 0;
    for (int k = 0; k < npanels; k++) {
        const int j1 = std::min(n, (k + 1) * kDenseNB);
        for (int j = k * kDenseNB; j < j1; j++) run = std::max(run, 3 * lastpose[j / 3] + 3);
        w.sky_panel_end[k] = std::max(run, j1);
    }
    int W = 1;
    double stored = 0.0;
    for (int c0 = 0; c0 < n; c0 += kDenseOuter) {
        const int cend = std::min(n, c0 + kDenseOuter);
        W = std::max(W, w.sky_panel_end[(cend - 1) / kDenseNB] - c0);
    }
    for (int j = 0; j < n; j++) stored += (double)(w.sky_panel_end[j / kDenseNB] - j);     // lower-triangle entries inside the row limits
    w.n = n; w.sky = true; w.sky_W = std::min(W + 1, n + 1);
    w.sky_fill = stored / (0.5 * (double)n * (n + 1));
}

template <typename S>
int ensure_skyline(bos_ctx* c) {
    if (c->sky_ready) return BOS_OK;
    Dev<S>& d = dev<S>(c);
    DenseWork<S>& w = swork<S>(c);
    skyline_symbolic<S>(c->P, w);
    w.Smat = c->mem.get<S>((size_t)w.n * w.sky_W);
    w.g = c->mem.get<S>((size_t)w.n);
    w.hllinv = c->mem.get<S>(3 * (size_t)std::max(d.NL, 1));
    w.ul = c->mem.get<S>(2 * (size_t)std::max(d.NL, 1));
    w.tl = c->mem.get<S>(2 * (size_t)std::max(d.NL, 1));
    w.tl_blk = c->mem.get<S>(64);
    if (!w.Smat || !w.g || !w.hllinv || !w.ul || !w.tl || !w.tl_blk) return fail(c, BOS_ERR_NOMEM, "skyline workspace allocation failed");
    c->sky_ready = true;
    return BOS_OK;
}
template <typename S>
int ensure_pcg(bos_ctx* c) {
    if (c->pcg_ready) return BOS_OK;
    Dev<S>& d = dev<S>(c);
    PcgWork<S>& w = pwork<S>(c);
    const size_t n = 3 * (size_t)d.NP, nl = (size_t)std::max(d.NL, 1);
    w.hllinv = c->mem.get<S>(3 * nl); w.ul = c->mem.get<S>(2 * nl); w.tl = c->mem.get<S>(2 * nl);
    w.Hlp = c->mem.get<S>(6 * (size_t)std::max(d.hpl_ld, 4));
    w.minv = c->mem.get<S>(6 * (size_t)d.NP);
    w.x = c->mem.get<S>(n); w.r = c->mem.get<S>(n); w.z = c->mem.get<S>(n);
    w.p0 = c->mem.get<S>(n); w.y = c->mem.get<S>(n);
    w.scal = c->mem.get<double>(32);
    w.bar = c->mem.get<unsigned>(4);
    w.Eb_pad = d.Eb_pad;
    {   // bearing omegas: one value for all edges is the common case (the reference never sets another, observation.hpp:17)
        const HostPattern& P = c->P;
        w.omega_uniform = 1; w.sqrt_omega = 1.0;
        const std::vector<double>& om = c->b_omega_sorted;
        if (!om.empty()) {
            for (double v : om) if (v != om[0]) { w.omega_uniform = 0; break; }
            w.sqrt_omega = std::sqrt(om[0]);
        }
        if (!w.omega_uniform) {
            std::vector<S> pw(P.pc_emap.size(), S(0));
            for (size_t k = 0; k < pw.size(); k++) if (P.pc_emap[k] >= 0) pw[k] = (S)std::sqrt(om[P.pc_emap[k]]);
            w.Pw = c->mem.upload(pw);
            std::vector<S> lw(P.ell_Lmap.size(), S(0));
            for (size_t k = 0; k < lw.size(); k++) if (P.ell_Lmap[k] >= 0) lw[k] = (S)std::sqrt(om[P.ell_Lmap[k]]);
            w.Lw = c->mem.upload(lw);
            std::vector<S> cw(P.lc_emap.size() + 32, S(0));
            for (size_t k = 0; k < P.lc_emap.size(); k++) if (P.lc_emap[k] >= 0) cw[k] = (S)std::sqrt(om[P.lc_emap[k]]);
            w.Cw = c->mem.upload(cw);
            if (!w.Pw || !w.Lw || !w.Cw) return fail(c, BOS_ERR_NOMEM, "pcg workspace allocation failed");
        }
    }
    w.hllinv_c = c->mem.get<S>(3 * (size_t)std::max(d.n_clm, 1));
    w.ul4 = c->mem.get<S>(4 * (size_t)std::max(d.n_clm, 1));
    w.z4 = c->mem.get<S>(8 * (size_t)d.NP);
    w.tpart = c->mem.get<S>(4 * (size_t)std::max(d.n_q, 1));
    w.qstat = c->mem.get<S>(5 * (size_t)std::max(d.n_q, 1));
    if (!w.tpart || !w.qstat) return fail(c, BOS_ERR_NOMEM, "pcg workspace allocation failed");
    w.rowS = c->mem.get<S>(24 * (size_t)d.pc_chunks * d.pc_cp);
    w.rS = c->mem.get<S>(3 * (size_t)d.pc_chunks * d.pc_cp);
    w.variant = c->opt.pcg_variant;
    w.precond = c->opt.pcg_precond;
    w.sm_count = c->sm_count;
    {   // chain preconditioner: factor storage in the solve's shared-memory layout (odd group stride: conflict-free columns)
        const size_t rows = (size_t)d.pc_chunks * d.pc_cp;
        w.ch_Kp = (d.pc_cp / 32) | 1;
        w.ch_cps = w.ch_Kp * 32;
        w.ch_fac_floats = 16 * w.ch_cps + 28 * w.ch_Kp;
        w.chD = c->mem.get<S>(6 * std::max<size_t>(rows, 1));
        w.chO = c->mem.get<S>(6 * std::max<size_t>(rows, 1));
        w.chF = c->mem.get<float>((size_t)std::max(d.pc_chunks, 1) * w.ch_fac_floats);
        if (!w.chD || !w.chO || !w.chF) return fail(c, BOS_ERR_NOMEM, "pcg workspace allocation failed");
        // coarse nodes: c_nseg segments of c_h = 32 m rows per chunk (pcg_coarse_nodes asked for; whole 32-row groups; at most 8)
        {
            const HostPattern& P = c->P;
            const int groups = std::max(d.pc_cp / 32, 1);
            int want = c->opt.pcg_coarse_nodes > 0 ? c->opt.pcg_coarse_nodes : 4;
            want = std::min(std::min(want, 8), groups);
            auto geometry = [&](int k) { const int m = (groups + k - 1) / k; w.c_h = 32 * m; w.c_nseg = (groups + m - 1) / m; };
            // half bandwidth of A_c in nodes: the span of the landmarks the assembly keeps (at most 16 nodes, k_coarse_lm) and of the pose-pose blocks
            auto band_nodes = [&]() {
                const int cp = d.pc_cp, h = w.c_h, ns = w.c_nseg;
                auto seg_of = [&](int pose) { const int cc = pose / cp; return cc * ns + (pose - cc * cp) / h; };
                int bwn = 1;
                for (int l = 0; l < P.NL; l++) {
                    int nn = 0, last = -2, first = -1;
                    bool over = false;
                    for (int q = P.lm_ptr[l]; q < P.lm_ptr[l + 1]; q++) {
                        const int gs = seg_of(P.lm_order_pose[q]);
                        if (gs == last) continue;
                        if (nn + 2 > 16) { over = true; break; }
                        nn += (last + 1 == gs) ? 1 : 2;
                        if (first < 0) first = gs;
                        last = gs;
                    }
                    if (!over && first >= 0) bwn = std::max(bwn, last + 1 - first);
                }
                for (size_t k = 0; k < P.off_lo.size(); k++) bwn = std::max(bwn, std::abs(seg_of(P.off_hi[k]) - seg_of(P.off_lo[k])) + 1);
                return bwn;
            };
            geometry(want);
            int bw = 3 * band_nodes() + 2;
            if (bw > 158 && w.c_nseg > 1) { geometry(1); bw = 3 * band_nodes() + 2; }   // wide coupling: fall back to one hat per chunk
            w.c_nc = 3 * (d.pc_chunks * w.c_nseg + 1);
            w.c_bw = (bw <= 158) ? std::min(bw, w.c_nc - 1) : 0;
            if (w.c_bw < 1 && w.c_nc > 1) w.c_bw = (bw <= 158) ? 1 : 0;
            w.coarse_refresh = c->opt.pcg_coarse_refresh > 0 ? c->opt.pcg_coarse_refresh : 8;
            w.coarse_valid = false; w.coarse_stale = false; w.coarse_age = 0; w.coarse_period = 0; w.coarse_its_last = 0; w.coarse_excess = 0.0;
        }
        w.cA = c->mem.get<double>((size_t)w.c_nc * w.c_nc);
        w.c_ld = (w.c_nc + 3) / 4 * 4;
        w.cAinv = c->mem.get<double>((size_t)w.c_nc * w.c_ld);
        w.cRc = c->mem.get<double>(6 * (size_t)std::max(d.pc_chunks * w.c_nseg, 1));
        w.cStats = c->mem.get<double>(8);
        w.cLc = c->mem.get<double>((size_t)w.c_nc * (w.c_bw + 1));
        w.cLr = c->mem.get<double>((size_t)w.c_nc * (w.c_bw + 1));
        w.cLdi = c->mem.get<double>((size_t)w.c_nc);
        if (!w.cA || !w.cAinv || !w.cRc || !w.cStats || !w.cLc || !w.cLr || !w.cLdi) return fail(c, BOS_ERR_NOMEM, "pcg workspace allocation failed");
        CUDA_OK(c, cudaMemsetAsync(w.cLr, 0, sizeof(double) * (size_t)w.c_nc * (w.c_bw + 1), c->stream));
    }
    w.xS = c->mem.get<S>(3 * (size_t)d.pc_chunks * d.pc_cp);
    if (!w.hllinv_c || !w.ul4 || !w.z4 || !w.rS || !w.xS || !w.rowS) return fail(c, BOS_ERR_NOMEM, "pcg workspace allocation failed");
    c->pcg_ready = true;
    return BOS_OK;
}

void close_peers(bos_ctx* c) {
    for (int r = 0; r < kMaxPeers; r++) {
        if (c->peer_opened[r]) cudaIpcCloseMemHandle(c->peer_opened[r]);
        c->peer_opened[r] = nullptr; c->peer_base[r] = nullptr;
    }
    c->peers_open = false;
    c->peer_epoch = 0;
    if (c->reduce_mode == 4 || c->reduce_mode == 5) c->reduce_mode = 3;
    c->dd.npeer = 0; c->df.npeer = 0; c->dd.peer_push = 0; c->df.peer_push = 0;
    c->peer_scratch = nullptr;   // owned by c->mem
    c->dd.stats_k2 = c->dd.stats; c->df.stats_k2 = c->df.stats;
}

// the peer window of rank r: values at peer_base[r], tail (statistics, barrier slots) tail_off bytes further (same problem => same layout)
template <typename S>
void apply_peer_mode(bos_ctx* c) {
    Dev<S>& d = dev<S>(c);
    const bool on = (c->reduce_mode == 4 || c->reduce_mode == 5) && c->peers_open && c->nranks > 1;
    d.npeer = on ? c->nranks : 0;
    d.peer_push = (on && c->reduce_mode == 4) ? 1 : 0;
    for (int r = 0; r < kMaxPeers; r++) {
        d.pv[r] = (on && r < c->nranks) ? reinterpret_cast<S*>(c->peer_base[r]) : nullptr;
        d.pstats[r] = (on && r < c->nranks) ? reinterpret_cast<double*>(c->peer_base[r] + c->tail_off) : nullptr;
    }
    d.stats_k2 = d.peer_push ? d.stats + kTailStatsK2 : d.stats;
}

int peer_barrier(bos_ctx* c, const double* publish) {
    PeerBarrier pb;
    for (int r = 0; r < kMaxPeers; r++) {
        unsigned char* tail = (r < c->nranks && c->peer_base[r]) ? c->peer_base[r] + c->tail_off : nullptr;
        pb.slots[r] = tail ? reinterpret_cast<unsigned long long*>(tail) + kTailSlots : nullptr;
        pb.pstats[r] = reinterpret_cast<double*>(tail);
    }
    pb.publish = publish;
    pb.error = reinterpret_cast<unsigned long long*>(c->peer_base[c->rank] + c->tail_off) + kTailError;
    pb.n = c->nranks; pb.rank = c->rank;
    pb.epoch = ++c->peer_epoch;
    return launch_peer_barrier(pb, c->stream);
}

int pick_solver(bos_ctx* c) {
    if (c->opt.solver == BOS_SOLVER_DENSE_CHOLESKY || c->opt.solver == BOS_SOLVER_PCG || c->opt.solver == BOS_SOLVER_SPARSE_CHOLESKY) return c->opt.solver;
    return (3 * c->P.NP <= c->opt.dense_max_dim) ? BOS_SOLVER_DENSE_CHOLESKY : BOS_SOLVER_PCG;
}

template <typename S>
int allreduce_impl(bos_ctx* c) {
    if (c->nranks > 1 && c->reduce_mode == 4) {
        // nothing left to move: the bearing kernel wrote every combined block into every replica.  The second cross-GPU barrier tells this
        // rank that ALL ranks' kernels have finished (their remote writes are visible after their kernel boundary) and publishes the
        // odometry share of the statistics.
        c->launches += peer_barrier(c, dev<S>(c).stats + kTailStatsK2);
        CUDA_OK(c, cudaGetLastError());
        return BOS_OK;
    }
    if (c->nranks > 1 && c->reduce_mode == 5) {
        // pull-based combine over peer memory: barrier (every rank's local build is complete) | pull the owners' pose ranges in place, sum the
        // landmark parts / rank-boundary poses / statistics of all ranks into a scratch | barrier (nobody reads this rank's parts any more) | commit
        Dev<S>& dp = dev<S>(c);
        PeerPull pp;
        pp.n = c->nranks; pp.rank = c->rank; pp.NP = c->P.NP; pp.NL = c->P.NL;
        for (int r = 0; r <= kMaxPeers; r++) pp.own_p0[r] = c->own_p0[std::min(r, c->nranks)];
        for (int r = 0; r < kMaxPeers; r++) {
            pp.bnd[r] = -1;
            if (r + 1 >= c->nranks || c->own_p0[r + 1] <= c->own_p0[r]) continue;
            const int pb = c->own_p0[r + 1] - 1;
            int64_t q[4];
            shard_ranges(c->P.Eb, c->P.Eo, r, c->nranks, q, nullptr);
            if (c->P.epose_ptr[pb + 1] > (int)q[1] && c->P.epose_ptr[pb] < (int)q[1]) pp.bnd[r] = pb;
        }
        c->launches += peer_barrier(c, nullptr);
        c->launches += launch_peer_pull<S>(dp, pp, static_cast<S*>(c->peer_scratch), c->sm_count, c->stream);
        c->launches += peer_barrier(c, nullptr);
        c->launches += launch_peer_commit<S>(dp, pp, static_cast<const S*>(c->peer_scratch), c->stream);
        CUDA_OK(c, cudaGetLastError());
        return BOS_OK;
    }
    if (c->nranks <= 1 || !c->comm) return BOS_OK;
    NcclApi& n = nccl();
    Dev<S>& d = dev<S>(c);
    const int dt = sizeof(S) == 8 ? kNcclFloat64 : kNcclFloat32;
    int rc;
    if (c->reduce_mode == 3 && c->P.slots_identity && n.GroupStart && n.GroupEnd) {
        // Ownership-based combine (reduce only what can overlap).  A pose block (b_p, Hpp) is complete on the rank whose tiles
        // own the pose, every rank computed every pose-pose block itself, the pose-landmark blocks stay rank-local (the fused PCG applies
        // that part of the operator from per-edge factors): only the landmark blocks and b_l are SUMMED (5 scalars per landmark), the
        // owned pose ranges are GATHERED (one broadcast per rank and array, grouped).
        const size_t NP = (size_t)d.NP, NL = (size_t)d.NL;
        rc = n.GroupStart();
        if (rc == 0 && NL > 0) rc = n.AllReduce(d.b + 3 * NP, d.b + 3 * NP, 2 * NL, dt, kNcclSum, c->comm, c->stream);
        if (rc == 0 && NL > 0) rc = n.AllReduce(d.Hll, d.Hll, 3 * NL, dt, kNcclSum, c->comm, c->stream);
        for (int r = 0; r < c->nranks && rc == 0; r++) {
            const size_t p0 = (size_t)c->own_p0[r];
            size_t cnt = (size_t)(c->own_p0[r + 1] - c->own_p0[r]);
            if (cnt == 0) continue;
            // the last owned pose's run may continue into the next rank's first tile: that one block is summed (k_hb_init zeroed it on every
            // rank, the owner holds damping + odometry + its part, the next rank the rest), everything else is gathered from its owner
            const int pb = c->own_p0[r + 1] - 1;
            int64_t q[4];
            shard_ranges(c->P.Eb, c->P.Eo, r, c->nranks, q, nullptr);
            if (r + 1 < c->nranks && c->P.epose_ptr[pb + 1] > (int)q[1] && c->P.epose_ptr[pb] < (int)q[1]) {
                rc = n.AllReduce(d.b + 3 * (size_t)pb, d.b + 3 * (size_t)pb, 3, dt, kNcclSum, c->comm, c->stream);
                if (rc == 0) rc = n.AllReduce(d.Hpp + 6 * (size_t)pb, d.Hpp + 6 * (size_t)pb, 6, dt, kNcclSum, c->comm, c->stream);
                cnt--;
                if (cnt == 0 || rc != 0) continue;
            }
            rc = n.Broadcast(d.b + 3 * p0, d.b + 3 * p0, 3 * cnt, dt, r, c->comm, c->stream);
            if (rc == 0) rc = n.Broadcast(d.Hpp + 6 * p0, d.Hpp + 6 * p0, 6 * cnt, dt, r, c->comm, c->stream);
        }
        const int rc2 = n.GroupEnd();
        if (rc == 0) rc = rc2;
    } else if (c->reduce_mode == 2 && c->P.slots_identity) {
        // b, diagonal blocks and pose-pose blocks only: all the fused PCG solve reads (it applies the pose-landmark part of the
        // operator from per-edge factors); the pose-landmark planes stay rank-local
        rc = n.AllReduce(d.vals, d.vals, c->vals_prefix, dt, kNcclSum, c->comm, c->stream);
    } else if (c->reduce_mode == 1 && c->P.slots_identity && c->nranks <= 8) {
        rc = n.AllReduce(d.vals, d.vals, c->vals_prefix, dt, kNcclSum, c->comm, c->stream);
        const size_t cnt = (size_t)c->shard_chunk_b;   // rank shards of every plane are equally sized (padded)
        for (int k = 0; k < 6 && rc == 0; k++) {
            S* plane = d.Hpl + (size_t)k * d.hpl_ld;
            rc = n.AllGather(plane + cnt * c->rank, plane, cnt, dt, c->comm, c->stream);
        }
    } else {
        rc = n.AllReduce(d.vals, d.vals, c->vals_len, dt, kNcclSum, c->comm, c->stream);
    }
    if (rc == 0) rc = n.AllReduce(d.stats, d.stats, 4, kNcclFloat64, kNcclSum, c->comm, c->stream);
    if (rc != 0) return fail(c, BOS_ERR_NCCL, std::string("nccl: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"));
    return BOS_OK;
}

template <typename S>
int linearize_impl(bos_ctx* c) {
    Dev<S>& d = dev<S>(c);
    const bool multi = c->nranks > 1;
    const bool zero_hpl = !c->P.slots_identity || (multi && !((c->reduce_mode >= 2 || (c->reduce_mode == 1 && c->nranks <= 8)) && c->P.slots_identity));
    const bool peer = multi && c->reduce_mode == 4, pull = multi && c->reduce_mode == 5;
    if ((peer || pull) && !(c->peers_open && c->P.slots_identity && c->nranks <= kMaxPeers))
        return fail(c, BOS_ERR_STATE, "reduce_mode 4 / 5 need bos_peer_open after the problem upload, at most 8 ranks and no duplicate (pose, landmark) edges");
    const bool owned = multi && (c->reduce_mode == 3 || peer || pull) && c->P.slots_identity;   // every rank writes every pose-pose block itself
    const bool zero_hoff = (multi && !owned) || c->P.has_shared_off;
    // NCCL modes SUM the replicas' landmark blocks: only rank 0 contributes the damping.  Peer mode: every replica receives every rank's REDs
    // on top of its own initialisation, so every replica starts from the full damping.
    const double damp_here = (c->rank == 0 || peer) ? c->opt.damping : 0.0;
    d.irls = c->robust_mode;
    if (peer) {
        apply_peer_mode<S>(c);
        // initialisation + odometry kernel (local replica only) | barrier: every replica is initialised | bearing kernel writing to all replicas
        c->launches += launch_linearize<S>(d, c->shard, c->opt.kernel_threshold, c->opt.damping, damp_here, zero_hpl, zero_hoff, c->sm_count, c->stream, false, c->rank, true, 1);
        c->launches += peer_barrier(c, nullptr);
        c->launches += launch_linearize<S>(d, c->shard, c->opt.kernel_threshold, c->opt.damping, damp_here, zero_hpl, zero_hoff, c->sm_count, c->stream, false, c->rank, true, 2);
    } else {
        if (d.npeer || pull) apply_peer_mode<S>(c);
        c->launches += launch_linearize<S>(d, c->shard, c->opt.kernel_threshold, c->opt.damping, damp_here, zero_hpl, zero_hoff, c->sm_count, c->stream, multi && !owned, c->rank, owned);
    }
    CUDA_OK(c, cudaGetLastError());
    c->linearized = true; c->solved = false;
    return BOS_OK;
}

template <typename S>
int solve_impl(bos_ctx* c) {
    Dev<S>& d = dev<S>(c);
    const int which = pick_solver(c);
    c->solver_used = which;
    if (c->nranks > 1 && c->reduce_mode >= 2 && c->P.slots_identity && !(which == BOS_SOLVER_PCG && c->opt.pcg_variant == 0))
        return fail(c, BOS_ERR_STATE, "reduce_mode 2 / 3 leave the pose-landmark blocks rank-local: only the fused PCG solve (pcg_variant 0) can follow");
    int nl = 0, rc = 0;
    if (which == BOS_SOLVER_DENSE_CHOLESKY || which == BOS_SOLVER_SPARSE_CHOLESKY) {
        const bool sky = which == BOS_SOLVER_SPARSE_CHOLESKY;
        int e = sky ? ensure_skyline<S>(c) : ensure_dense<S>(c);
        if (e) return e;
        rc = launch_dense_solve<S>(d, sky ? swork<S>(c) : dwork<S>(c), c->opt.damping, c->stream, &nl);
        c->stats.pcg_iterations = 0;
        c->stats.precond_used = -1; c->stats.pcg_resolves = 0;
        c->pcg_bad = false; c->pcg_capped = false;
    } else {
        int e = ensure_pcg<S>(c);
        if (e) return e;
        int iters = 0;
        double rtol = c->opt.pcg_rtol;
        if (sizeof(S) == 4 && rtol < 1e-6) rtol = 1e-6;
        // IRLS re-weights every edge every iteration; the fused kernel re-derives its per-edge factors from the state and the STATIC omegas, so
        // IRLS solves take the classic loop, which applies the stored pose-landmark blocks
        pwork<S>(c).variant = c->robust_mode ? 1 : c->opt.pcg_variant;
        pwork<S>(c).host_scal = c->pinned;
        rc = launch_pcg_solve<S>(d, pwork<S>(c), c->opt.pcg_max_iters, rtol, c->stream, &iters, &nl);
        if (rc < 0) return fail(c, BOS_ERR_CUDA, std::string("pcg: ") + cudaGetErrorString(cudaGetLastError()));
        c->stats.pcg_iterations = iters;
        c->stats.precond_used = pwork<S>(c).precond_used; c->stats.pcg_resolves = pwork<S>(c).resolves;
        c->pcg_bad = (rc == 1);
        c->pcg_capped = (rc == 0 && iters >= c->opt.pcg_max_iters);
    }
    c->launches += nl;
    CUDA_OK(c, cudaGetLastError());
    if (c->nranks > 1 && c->comm) {
        // Replicated solves are not bitwise identical (atomic summation orders differ), and every rank linearizes its edge shard at
        // ITS state: rank 0's increment is the one everybody applies, so the replicas never drift apart.  (The implicit Schur
        // complement cancels terms ~1e8 times larger than its small eigenvalues: a 1e-8 inconsistency between the shards' states
        // is enough to cost it positive definiteness.)
        NcclApi& n = nccl();
        const int dt = (sizeof(S) == 8) ? kNcclFloat64 : kNcclFloat32;
        if (n.Broadcast(d.delta, d.delta, (size_t)(3 * (size_t)d.NP + 2 * (size_t)d.NL), dt, 0, c->comm, c->stream) != 0)
            return fail(c, BOS_ERR_NCCL, "ncclBroadcast of the increment failed");
    }
    c->solved = true; c->delta_valid = true;
    return BOS_OK;
}

template <typename S>
int update_impl(bos_ctx* c) {
    c->launches += launch_update<S>(dev<S>(c), c->stream);
    CUDA_OK(c, cudaGetLastError());
    c->linearized = false; c->solved = false; c->delta_valid = false;
    return BOS_OK;
}

template <typename S>
int fetch_stats(bos_ctx* c) {
    double h_pageable[kTailError + 1];
    double* h = c->pinned ? c->pinned + 32 : h_pageable;
    CUDA_OK(c, cudaMemcpyAsync(h, dev<S>(c).stats, sizeof(h_pageable), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    unsigned long long peer_err;
    std::memcpy(&peer_err, &h[kTailError], sizeof(peer_err));
    if (peer_err != 0) return fail(c, BOS_ERR_NCCL, "reduce_mode 4: a cross-GPU barrier timed out (a rank did not reach bos_linearize)");
    c->stats.chi2_bearing = h[0]; c->stats.chi2_odometry = h[1];
    c->stats.over_bearing = (int64_t)llround(h[2]); c->stats.over_odometry = (int64_t)llround(h[3]);
    c->stats.delta_inf = h[4];
    c->stats.state_digest = h[6];
    c->stats.solver_status = (h[5] != 0.0 || c->pcg_bad) ? 1 : (c->pcg_capped ? 2 : 0);
    c->stats.solver_used = c->solver_used;
    return BOS_OK;
}

template <typename S>
int step_impl(bos_ctx* c) {
    c->launches = 0;
    int rc;
    cudaEventRecord(c->ev[0], c->stream);
    if ((rc = linearize_impl<S>(c))) return rc;
    cudaEventRecord(c->ev[1], c->stream);
    if ((rc = allreduce_impl<S>(c))) return rc;
    cudaEventRecord(c->ev[2], c->stream);
    if ((rc = solve_impl<S>(c))) return rc;
    cudaEventRecord(c->ev[3], c->stream);
    if ((rc = update_impl<S>(c))) return rc;
    cudaEventRecord(c->ev[4], c->stream);
    if ((rc = fetch_stats<S>(c))) return rc;
    cudaEventElapsedTime(&c->stats.ms_linearize, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->stats.ms_allreduce, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&c->stats.ms_solve, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&c->stats.ms_update, c->ev[3], c->ev[4]);
    c->stats.gpu_launches = c->launches;
    return BOS_OK;
}

template <typename S>
int step_lm_impl(bos_ctx* c, double* chi2_after, int* accepted, double* damping_next) {
    Dev<S>& d = dev<S>(c);
    const size_t pb = 4 * (size_t)d.NP * sizeof(S), lb = 2 * (size_t)std::max(d.NL, 1) * sizeof(S);
    if (!c->lm_pose_bak) {
        c->lm_pose_bak = c->mem.get<unsigned char>(pb);
        c->lm_lm_bak = c->mem.get<unsigned char>(lb);
        if (!c->lm_pose_bak || !c->lm_lm_bak) return fail(c, BOS_ERR_NOMEM, "LM backup allocation failed");
    }
    CUDA_OK(c, cudaMemcpyAsync(c->lm_pose_bak, d.pose, pb, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->lm_lm_bak, d.lm, lb, cudaMemcpyDeviceToDevice, c->stream));
    int rc = step_impl<S>(c);                       // chi2 of the state BEFORE the update is in c->stats
    if (rc) return rc;
    const bos_stats taken = c->stats;
    const double before = taken.chi2_bearing + taken.chi2_odometry;
    if ((rc = linearize_impl<S>(c))) return rc;     // chi2 at the new state
    if ((rc = allreduce_impl<S>(c))) return rc;
    if ((rc = fetch_stats<S>(c))) return rc;
    const double after = c->stats.chi2_bearing + c->stats.chi2_odometry;
    const bool ok = after < before;
    if (!ok) {
        CUDA_OK(c, cudaMemcpyAsync(d.pose, c->lm_pose_bak, pb, cudaMemcpyDeviceToDevice, c->stream));
        CUDA_OK(c, cudaMemcpyAsync(d.lm, c->lm_lm_bak, lb, cudaMemcpyDeviceToDevice, c->stream));
        launch_pose_theta<S>(d, c->stream);
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        c->linearized = false;
        c->opt.damping = std::min(c->opt.damping * 10.0, 1e9);
    } else {
        c->opt.damping = std::max(c->opt.damping / 3.0, 1e-9);
    }
    c->stats = taken;
    if (chi2_after) *chi2_after = after;
    if (accepted) *accepted = ok ? 1 : 0;
    if (damping_next) *damping_next = c->opt.damping;
    return BOS_OK;
}

template <typename S>
int set_state_impl(bos_ctx* c, const double* poses, const double* lms) {
    Dev<S>& d = dev<S>(c);
    if (poses) {
        if (sizeof(S) == 8) CUDA_OK(c, cudaMemcpyAsync(d.pose, poses, 4 * (size_t)d.NP * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        else {
            std::vector<S> v = narrow<S>(poses, 4 * (size_t)d.NP);
            CUDA_OK(c, cudaMemcpyAsync(d.pose, v.data(), v.size() * sizeof(S), cudaMemcpyHostToDevice, c->stream));
            CUDA_OK(c, cudaStreamSynchronize(c->stream));
        }
    }
    if (poses) { launch_pose_theta<S>(d, c->stream); CUDA_OK(c, cudaGetLastError()); }
    if (lms && d.NL > 0) {
        if (sizeof(S) == 8) CUDA_OK(c, cudaMemcpyAsync(d.lm, lms, 2 * (size_t)d.NL * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        else {
            std::vector<S> v = narrow<S>(lms, 2 * (size_t)d.NL);
            CUDA_OK(c, cudaMemcpyAsync(d.lm, v.data(), v.size() * sizeof(S), cudaMemcpyHostToDevice, c->stream));
            CUDA_OK(c, cudaStreamSynchronize(c->stream));
        }
    }
    c->linearized = false; c->solved = false;
    if (!c->stepping_host) { c->pwd.coarse_valid = false; c->pwf.coarse_valid = false; }   // a state from outside: the cached coarse operator describes another one
    return BOS_OK;
}

template <typename S>
int get_array(bos_ctx* c, const S* dptr, double* out, size_t n) {
    if (!out || n == 0) return BOS_OK;
    if (sizeof(S) == 8) {
        CUDA_OK(c, cudaMemcpyAsync(out, dptr, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
    } else {
        std::vector<S> v(n);
        CUDA_OK(c, cudaMemcpyAsync(v.data(), dptr, n * sizeof(S), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < n; i++) out[i] = (double)v[i];
    }
    return BOS_OK;
}

// expanded (full) block arrays from the compact device layout
template <typename S>
int download_blocks_impl(bos_ctx* c, double* Hpp, double* Hll, double* Hpl, double* Hoff, double* b) {
    Dev<S>& d = dev<S>(c);
    std::vector<double> vals(c->vals_len);
    int rc = get_array<S>(c, d.vals, vals.data(), c->vals_len);
    if (rc) return rc;
    const double* vb = vals.data();
    const double* vpp = vb + d.N;
    const double* vll = vpp + 6 * (size_t)d.NP;
    const double* voff = vll + 3 * (size_t)d.NL;
    const double* vpl = vals.data() + c->vals_prefix;
    if (b) std::copy(vb, vb + d.N, b);
    if (Hpp)
        for (int p = 0; p < d.NP; p++) {
            const double* h = vpp + 6 * (size_t)p;
            const double m[9] = {h[0], h[1], h[2], h[1], h[3], h[4], h[2], h[4], h[5]};
            std::copy(m, m + 9, Hpp + 9 * (size_t)p);
        }
    if (Hll)
        for (int l = 0; l < d.NL; l++) {
            const double* h = vll + 3 * (size_t)l;
            const double m[4] = {h[0], h[1], h[1], h[2]};
            std::copy(m, m + 4, Hll + 4 * (size_t)l);
        }
    if (Hoff) std::copy(voff, voff + 9 * (size_t)d.n_off, Hoff);
    if (Hpl)
        for (int s = 0; s < d.n_hpl; s++)
            for (int k = 0; k < 6; k++) Hpl[6 * (size_t)s + k] = vpl[(size_t)k * d.hpl_ld + s];
    return BOS_OK;
}

template <typename S>
int edge_terms_impl(bos_ctx* c, double* err_b, double* jac_b, double* err_o, double* jac_o) {
    Dev<S>& d = dev<S>(c);
    DevAlloc tmp;
    S* eb = tmp.get<S>((size_t)d.Eb); S* jb = tmp.get<S>(5 * (size_t)d.Eb);
    S* eo = tmp.get<S>(3 * (size_t)d.Eo); S* jo = tmp.get<S>(18 * (size_t)d.Eo);
    if (!eb || !jb || !eo || !jo) return fail(c, BOS_ERR_NOMEM, "edge term scratch allocation failed");
    launch_edge_terms<S>(d, eb, jb, eo, jo, c->stream);
    CUDA_OK(c, cudaGetLastError());
    int rc;
    if ((rc = get_array<S>(c, eb, err_b, (size_t)d.Eb))) return rc;
    if ((rc = get_array<S>(c, jb, jac_b, 5 * (size_t)d.Eb))) return rc;
    if ((rc = get_array<S>(c, eo, err_o, 3 * (size_t)d.Eo))) return rc;
    if ((rc = get_array<S>(c, jo, jac_o, 18 * (size_t)d.Eo))) return rc;
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BOS_OK;
}

template <typename S>
int triangulate_impl(bos_ctx* c, int* single) {
    Dev<S>& d = dev<S>(c);
    CUDA_OK(c, cudaMemsetAsync(c->d_single_obs, 0, sizeof(int), c->stream));
    launch_triangulate<S>(d, c->d_single_obs, c->stream);
    CUDA_OK(c, cudaGetLastError());
    int h = 0;
    CUDA_OK(c, cudaMemcpyAsync(&h, c->d_single_obs, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    if (single) *single = h;
    return BOS_OK;
}

#define DISPATCH(c, fn, ...) ((c)->f64() ? fn<double>(__VA_ARGS__) : fn<float>(__VA_ARGS__))
#define NEED(c, cond, msg) \
    if (!(cond)) return fail(c, BOS_ERR_STATE, msg)

}  // namespace

extern "C" {

void bos_default_options(bos_options* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->device = 0;
    o->precision = BOS_PRECISION_F64;
    o->solver = BOS_SOLVER_AUTO;
    o->dense_max_dim = 192;   // measured crossover (profiles/solver_sweep_r02.jsonl): the PCG wins from 3 NP ~ 300 on
    o->kernel_threshold = 1.0;
    o->damping = 0.01f;  // the reference's float literal, widened (slam/solver.cpp:17)
    o->pcg_max_iters = 20000;
    o->pcg_rtol = 1e-10;
}

int bos_version(void) { return 100; }

int bos_create(const bos_options* opts, bos_ctx** out) {
    if (!out) return BOS_ERR_INVALID;
    *out = nullptr;
    std::unique_ptr<bos_ctx> c(new bos_ctx());
    if (opts) c->opt = *opts; else bos_default_options(&c->opt);
    std::memset(&c->stats, 0, sizeof(c->stats));
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BOS_ERR_CUDA;
    if (c->opt.device < 0 || c->opt.device >= ndev) return BOS_ERR_INVALID;
    if (cudaSetDevice(c->opt.device) != cudaSuccess) return BOS_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->opt.device) != cudaSuccess) return BOS_ERR_CUDA;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return BOS_ERR_CUDA;
    if (cudaHostAlloc(reinterpret_cast<void**>(&c->pinned), 64 * sizeof(double), cudaHostAllocDefault) != cudaSuccess) { c->pinned = nullptr; cudaGetLastError(); }
    for (auto& e : c->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return BOS_ERR_CUDA;
    *out = c.release();
    return BOS_OK;
}

int bos_destroy(bos_ctx* c) {
    if (!c) return BOS_OK;
    cudaSetDevice(c->opt.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    close_peers(c);
    for (cudaGraphExec_t* g : {&c->dwd.graph, &c->dwf.graph, &c->swd.graph, &c->swf.graph})
        if (*g) { cudaGraphExecDestroy(*g); *g = nullptr; }
    for (int k = 0; k < 3; k++) {
        if (c->pwd.coarse_ev[k]) cudaEventDestroy(c->pwd.coarse_ev[k]);
        if (c->pwf.coarse_ev[k]) cudaEventDestroy(c->pwf.coarse_ev[k]);
    }
    c->mem.release();
    if (c->pinned) cudaFreeHost(c->pinned);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BOS_OK;
}

const char* bos_last_error(const bos_ctx* c) { return c ? c->err.c_str() : "null context"; }

int bos_set_kernel_threshold(bos_ctx* c, double kt) {
    if (!c) return BOS_ERR_INVALID;
    c->opt.kernel_threshold = kt;
    return BOS_OK;
}
int bos_set_robust_mode(bos_ctx* c, int mode) {
    if (!c || (mode != BOS_ROBUST_REFERENCE && mode != BOS_ROBUST_IRLS)) return BOS_ERR_INVALID;
    if (mode == BOS_ROBUST_IRLS && c->nranks > 1 && c->reduce_mode >= 2)
        return fail(c, BOS_ERR_STATE, "IRLS needs the stored pose-landmark blocks: use reduce_mode 0 or 1 with several ranks");
    c->robust_mode = mode;
    return BOS_OK;
}
int bos_set_damping_factor(bos_ctx* c, double df) {
    if (!c) return BOS_ERR_INVALID;
    c->opt.damping = df;
    return BOS_OK;
}

int bos_upload_problem(bos_ctx* c, int NP, int NL, int fixed_pose_stix, int64_t Eb, const int32_t* b_pose, const int32_t* b_lm,
                       const double* b_z, const double* b_omega, int64_t Eo, const int32_t* o_src, const int32_t* o_dst,
                       const double* o_z, const double* o_omega) {
    if (!c) return BOS_ERR_INVALID;
    if ((Eb > 0 && (!b_pose || !b_lm || !b_z)) || (Eo > 0 && (!o_src || !o_dst || !o_z || !o_omega)))
        return fail(c, BOS_ERR_INVALID, "null edge array");
    // the kernels keep the upper triangle of every odometry Omega: refuse a matrix the reference would treat differently
    // (it multiplies with the full 3x3, slam/solver.cpp:60-61; utils/g2o_utils.cpp:91-106 always builds a symmetric one)
    for (int64_t e = 0; e < Eo; e++) {
        const double* om = o_omega + 9 * e;
        double mx = 0.0;
        for (int k = 0; k < 9; k++) mx = std::max(mx, std::fabs(om[k]));
        if (std::fabs(om[1] - om[3]) > 1e-12 * mx || std::fabs(om[2] - om[6]) > 1e-12 * mx || std::fabs(om[5] - om[7]) > 1e-12 * mx)
            return fail(c, BOS_ERR_INVALID, "odometry omega of edge " + std::to_string(e) + " is not symmetric");
    }
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    c->have_problem = false; c->delta_valid = false; c->dense_ready = false; c->pcg_ready = false; c->sky_ready = false;
    c->lm_pose_bak = nullptr; c->lm_lm_bak = nullptr;
    close_peers(c);
    c->mem.release();
    PatternCore core;
    const auto t0 = std::chrono::steady_clock::now();
    // auto: the GPU sorts pay from ~200 k edges on (mini / full build in microseconds on the host); the tables are identical either way
    bool on_device = c->device_setup == 1 || (c->device_setup < 0 && Eb >= 200000);
    if (on_device) {
        std::string err;
        const int rcd = device_pattern_core(core, NP, NL, Eb, b_pose, b_lm, c->stream, err);
        if (rcd == 1) return fail(c, BOS_ERR_INVALID, err);
        if (rcd != 0) {
            if (c->device_setup == 1) return fail(c, BOS_ERR_CUDA, err);
            cudaGetLastError();
            core = PatternCore();
            on_device = false;     // auto: no room for the sort buffers -- the host builder produces the same tables
        }
    }
    c->device_setup_used = on_device;
    const auto t1 = std::chrono::steady_clock::now();
    if (build_pattern(c->P, NP, NL, fixed_pose_stix, Eb, b_pose, b_lm, Eo, o_src, o_dst, c->sm_count, on_device ? &core : nullptr) != 0)
        return fail(c, BOS_ERR_INVALID, c->P.error);
    c->setup_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    c->setup_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    int rc = c->f64() ? upload_impl<double>(c, b_z, b_omega, o_z, o_omega) : upload_impl<float>(c, b_z, b_omega, o_z, o_omega);
    if (rc) return rc;
    compute_shard(c);
    c->have_problem = true; c->linearized = false; c->solved = false;
    return BOS_OK;
}

int bos_set_state(bos_ctx* c, const double* poses, const double* lms) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "set_state before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    int rc = DISPATCH(c, set_state_impl, c, poses, lms);
    if (rc) return rc;
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BOS_OK;
}

int bos_get_state(bos_ctx* c, double* poses, double* lms) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "get_state before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    int rc;
    if (c->f64()) {
        if ((rc = get_array<double>(c, c->dd.pose, poses, 4 * (size_t)c->dd.NP))) return rc;
        return get_array<double>(c, c->dd.lm, lms, 2 * (size_t)c->dd.NL);
    }
    if ((rc = get_array<float>(c, c->df.pose, poses, 4 * (size_t)c->df.NP))) return rc;
    return get_array<float>(c, c->df.lm, lms, 2 * (size_t)c->df.NL);
}

int bos_linearize(bos_ctx* c) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "linearize before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    c->launches = 0;
    cudaEventRecord(c->ev[0], c->stream);
    int rc = DISPATCH(c, linearize_impl, c);
    if (rc) return rc;
    cudaEventRecord(c->ev[1], c->stream);
    if ((rc = DISPATCH(c, allreduce_impl, c))) return rc;
    cudaEventRecord(c->ev[2], c->stream);
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&c->stats.ms_linearize, c->ev[0], c->ev[1]);   // readable through bos_get_stats
    cudaEventElapsedTime(&c->stats.ms_allreduce, c->ev[1], c->ev[2]);
    c->stats.gpu_launches = c->launches;
    return BOS_OK;
}

int bos_solve(bos_ctx* c) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem && c->linearized, "solve before linearize");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    int rc = DISPATCH(c, solve_impl, c);
    if (rc) return rc;
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BOS_OK;
}

int bos_update(bos_ctx* c) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "update before upload_problem");
    NEED(c, c->delta_valid, "update without an increment: call bos_solve or bos_upload_delta first (an increment is applied once)");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    int rc = DISPATCH(c, update_impl, c);
    if (rc) return rc;
    return DISPATCH(c, fetch_stats, c);
}

int bos_step(bos_ctx* c, bos_stats* stats) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "step before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    int rc = DISPATCH(c, step_impl, c);
    if (rc) return rc;
    if (stats) *stats = c->stats;
    return BOS_OK;
}

int bos_step_host(bos_ctx* c, double* poses, double* lms, bos_stats* stats) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "step before upload_problem");
    if (!poses) return fail(c, BOS_ERR_INVALID, "null state");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    c->stepping_host = true;
    int rc = DISPATCH(c, set_state_impl, c, poses, lms);
    c->stepping_host = false;
    if (rc) return rc;
    if ((rc = DISPATCH(c, step_impl, c))) return rc;
    if ((rc = bos_get_state(c, poses, lms))) return rc;
    if (stats) *stats = c->stats;
    return BOS_OK;
}

int bos_step_lm(bos_ctx* c, bos_stats* stats, double* chi2_after, int* accepted, double* damping_next) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "step before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    int rc = DISPATCH(c, step_lm_impl, c, chi2_after, accepted, damping_next);
    if (rc) return rc;
    if (stats) *stats = c->stats;
    return BOS_OK;
}

int bos_get_stats(bos_ctx* c, bos_stats* stats) {
    if (!c || !stats) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "get_stats before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    int rc = DISPATCH(c, fetch_stats, c);
    if (rc) return rc;
    *stats = c->stats;
    return BOS_OK;
}

int bos_triangulate(bos_ctx* c, int* single_obs_count) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "triangulate before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    return DISPATCH(c, triangulate_impl, c, single_obs_count);
}

int bos_pattern_info_get(bos_ctx* c, bos_pattern_info* out) {
    if (!c || !out) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "pattern before upload_problem");
    out->n_hpl = (int64_t)c->P.slot_pose.size();
    out->n_hpp_off = (int64_t)c->P.off_lo.size();
    build_csc(c->P);
    out->csc_n = c->P.N - 3;
    out->csc_nnz = (int64_t)c->P.csc_rowidx.size();
    out->N = c->P.N;
    out->vals_len = (int64_t)c->vals_len;
    return BOS_OK;
}

int bos_download_pattern(bos_ctx* c, int32_t* hpl_pose, int32_t* hpl_lm, int32_t* off_lo, int32_t* off_hi, int64_t* b_slot, int64_t* o_slot) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "pattern before upload_problem");
    const HostPattern& P = c->P;
    if (hpl_pose) std::copy(P.slot_pose.begin(), P.slot_pose.end(), hpl_pose);
    if (hpl_lm) std::copy(P.slot_lm.begin(), P.slot_lm.end(), hpl_lm);
    if (off_lo) std::copy(P.off_lo.begin(), P.off_lo.end(), off_lo);
    if (off_hi) std::copy(P.off_hi.begin(), P.off_hi.end(), off_hi);
    if (b_slot)
        for (int k = 0; k < P.Eb; k++) b_slot[P.b_perm[k]] = P.b_slot[k];  // caller's edge order
    if (o_slot)
        for (int e = 0; e < P.Eo; e++) o_slot[e] = P.o_slot[e];
    return BOS_OK;
}

int bos_download_blocks(bos_ctx* c, double* Hpp, double* Hll, double* Hpl, double* Hoff, double* b) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem && c->linearized, "download_blocks before linearize");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    return DISPATCH(c, download_blocks_impl, c, Hpp, Hll, Hpl, Hoff, b);
}

int bos_download_csc(bos_ctx* c, int32_t* colptr, int32_t* rowidx, double* val, double* b_nofixed) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "download_csc before upload_problem");
    build_csc(c->P);
    const HostPattern& P = c->P;
    if (colptr) std::copy(P.csc_colptr.begin(), P.csc_colptr.end(), colptr);
    if (rowidx) std::copy(P.csc_rowidx.begin(), P.csc_rowidx.end(), rowidx);
    if (!val && !b_nofixed) return BOS_OK;
    NEED(c, c->linearized, "download_csc values before linearize");
    std::vector<double> Hpp(9 * (size_t)P.NP), Hll(4 * (size_t)std::max(P.NL, 1)), Hpl(6 * P.slot_pose.size() + 1), Hoff(9 * P.off_lo.size() + 1), b(P.N);
    int rc = bos_download_blocks(c, Hpp.data(), Hll.data(), Hpl.data(), Hoff.data(), b.data());
    if (rc) return rc;
    if (val) {
        const double* src[4] = {Hpp.data(), Hll.data(), Hoff.data(), Hpl.data()};
        for (size_t k = 0; k < P.csc_rowidx.size(); k++) val[k] = src[P.csc_src_kind[k]][P.csc_src_index[k]];
    }
    if (b_nofixed) {
        const int f3 = 3 * P.fixed;
        for (int i = 0, k = 0; i < P.N; i++)
            if (i < f3 || i >= f3 + 3) b_nofixed[k++] = b[i];
    }
    return BOS_OK;
}

int bos_download_delta(bos_ctx* c, double* delta) {
    if (!c || !delta) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "download_delta before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    if (c->f64()) return get_array<double>(c, c->dd.delta, delta, (size_t)c->P.N);
    return get_array<float>(c, c->df.delta, delta, (size_t)c->P.N);
}

int bos_upload_delta(bos_ctx* c, const double* delta) {
    if (!c || !delta) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "upload_delta before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    if (c->f64()) {
        CUDA_OK(c, cudaMemcpy(c->dd.delta, delta, (size_t)c->P.N * sizeof(double), cudaMemcpyHostToDevice));
    } else {
        std::vector<float> v = narrow<float>(delta, (size_t)c->P.N);
        CUDA_OK(c, cudaMemcpy(c->df.delta, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    c->delta_valid = true;
    return BOS_OK;
}

int bos_edge_terms(bos_ctx* c, double* err_b, double* jac_b, double* err_o, double* jac_o) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "edge_terms before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    return DISPATCH(c, edge_terms_impl, c, err_b, jac_b, err_o, jac_o);
}

int bos_triangulate_landmarks(const bos_options* opts, int NP, const double* poses, int64_t Eb, const int32_t* b_pose, const int32_t* b_lm,
                              const double* b_z, int NL, double* lms, int* single_obs_count) {
    if (!poses || !lms || NP <= 0 || NL < 0 || Eb < 0) return BOS_ERR_INVALID;
    bos_ctx* c = nullptr;
    int rc = bos_create(opts, &c);
    if (rc) return rc;
    rc = bos_upload_problem(c, NP, NL, 0, Eb, b_pose, b_lm, b_z, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    if (!rc) rc = bos_set_state(c, poses, nullptr);
    if (!rc) rc = bos_triangulate(c, single_obs_count);
    if (!rc) rc = bos_get_state(c, nullptr, lms);
    bos_destroy(c);
    return rc;
}

int bos_eval_bearing_edges(const bos_options* opts, int64_t n, const double* poses, const double* lms, const double* z, double* err, double* jac5) {
    if (n <= 0 || n > 0x3fffffff || !poses || !lms || !z || !err || !jac5) return BOS_ERR_INVALID;
    std::vector<int32_t> idx((size_t)n);
    for (int64_t i = 0; i < n; i++) idx[(size_t)i] = (int32_t)i;
    bos_ctx* c = nullptr;
    int rc = bos_create(opts, &c);
    if (rc) return rc;
    rc = bos_upload_problem(c, (int)n, (int)n, 0, n, idx.data(), idx.data(), z, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    if (!rc) rc = bos_set_state(c, poses, lms);
    if (!rc) rc = bos_edge_terms(c, err, jac5, nullptr, nullptr);
    bos_destroy(c);
    return rc;
}

int bos_eval_odometry_edges(const bos_options* opts, int64_t n, const double* src, const double* dst, const double* z3, double* err3, double* jac18) {
    if (n <= 0 || n > 0x1fffffff || !src || !dst || !z3 || !err3 || !jac18) return BOS_ERR_INVALID;
    std::vector<int32_t> s((size_t)n), t((size_t)n);
    std::vector<double> poses(8 * (size_t)n), om(9 * (size_t)n, 0.0);
    for (int64_t i = 0; i < n; i++) {
        s[(size_t)i] = (int32_t)i; t[(size_t)i] = (int32_t)(n + i);
        std::copy(src + 4 * i, src + 4 * i + 4, poses.begin() + 4 * i);
        std::copy(dst + 4 * i, dst + 4 * i + 4, poses.begin() + 4 * (n + i));
        om[9 * (size_t)i] = om[9 * (size_t)i + 4] = om[9 * (size_t)i + 8] = 1.0;
    }
    bos_ctx* c = nullptr;
    int rc = bos_create(opts, &c);
    if (rc) return rc;
    rc = bos_upload_problem(c, (int)(2 * n), 0, 0, 0, nullptr, nullptr, nullptr, nullptr, n, s.data(), t.data(), z3, om.data());
    if (!rc) rc = bos_set_state(c, poses.data(), nullptr);
    if (!rc) rc = bos_edge_terms(c, nullptr, nullptr, err3, jac18);
    bos_destroy(c);
    return rc;
}

int bos_nccl_unique_id(char* uid128) {
    if (!uid128) return BOS_ERR_INVALID;
    NcclApi& n = nccl();
    if (!n.ok) return BOS_ERR_NCCL;
    nccl_uid_t id;
    if (n.GetUniqueId(&id) != 0) return BOS_ERR_NCCL;
    std::memcpy(uid128, id.internal, BOS_NCCL_UID_BYTES);
    return BOS_OK;
}

int bos_comm_init(bos_ctx* c, int rank, int nranks, const char* uid128) {
    if (!c || !uid128 || nranks < 1 || rank < 0 || rank >= nranks) return BOS_ERR_INVALID;
    if (nranks > 64) return fail(c, BOS_ERR_INVALID, "at most 64 ranks");
    NcclApi& n = nccl();
    if (!n.ok) return fail(c, BOS_ERR_NCCL, "libnccl.so.2 not loadable");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    nccl_uid_t id;
    std::memcpy(id.internal, uid128, BOS_NCCL_UID_BYTES);
    int rc = n.CommInitRank(&c->comm, nranks, id, rank);
    if (rc != 0) return fail(c, BOS_ERR_NCCL, std::string("ncclCommInitRank: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"));
    c->rank = rank; c->nranks = nranks;
    if (c->have_problem) compute_shard(c);
    return BOS_OK;
}

int bos_set_reduce_mode(bos_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 5) return BOS_ERR_INVALID;
    if (mode >= 4 && !c->peers_open) return fail(c, BOS_ERR_STATE, "reduce_mode 4 / 5: call bos_peer_open first (after bos_upload_problem)");
    c->reduce_mode = mode;
    return BOS_OK;
}

// ---- reduce_mode 4: the value buffers of the ranks of one box, mapped into each other's address space (CUDA IPC over NVLink) ----
int bos_peer_export(bos_ctx* c, void* handle64, int64_t* offset) {
    if (!c || !handle64 || !offset) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "peer export before upload_problem");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    unsigned char* vals = c->f64() ? reinterpret_cast<unsigned char*>(c->dd.vals) : reinterpret_cast<unsigned char*>(c->df.vals);
    // small cudaMalloc blocks are sub-allocated: the IPC handle names the whole underlying allocation, so the offset inside it travels along
    typedef int (*get_range_t)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    unsigned long long base = reinterpret_cast<unsigned long long>(vals);
    size_t size = 0;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn && qr == cudaDriverEntryPointSuccess) {
        if (reinterpret_cast<get_range_t>(fn)(&base, &size, reinterpret_cast<unsigned long long>(vals)) != 0) base = reinterpret_cast<unsigned long long>(vals);
    } else {
        cudaGetLastError();
    }
    cudaIpcMemHandle_t h;
    CUDA_OK(c, cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
    static_assert(sizeof(cudaIpcMemHandle_t) == BOS_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    std::memcpy(handle64, &h, sizeof(h));
    *offset = (int64_t)(reinterpret_cast<unsigned long long>(vals) - base);
    return BOS_OK;
}

int bos_peer_open(bos_ctx* c, const void* handles, const int64_t* offsets) {
    if (!c || !handles || !offsets) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "peer open before upload_problem");
    if (c->nranks < 2 || c->nranks > kMaxPeers) return fail(c, BOS_ERR_STATE, "bos_peer_open: 2 to 8 ranks (bos_comm_init / bos_set_edge_shard first)");
    CUDA_OK(c, cudaSetDevice(c->opt.device));
    close_peers(c);
    unsigned char* own = c->f64() ? reinterpret_cast<unsigned char*>(c->dd.vals) : reinterpret_cast<unsigned char*>(c->df.vals);
    for (int r = 0; r < c->nranks; r++) {
        if (r == c->rank) { c->peer_base[r] = own; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * BOS_IPC_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            const std::string why = cudaGetErrorString(cudaGetLastError());
            close_peers(c);
            return fail(c, BOS_ERR_CUDA, "cudaIpcOpenMemHandle(rank " + std::to_string(r) + "): " + why);
        }
        c->peer_opened[r] = p;
        c->peer_base[r] = static_cast<unsigned char*>(p) + offsets[r];
    }
    c->peer_scratch = c->mem.get<double>((size_t)peer_scratch_stats_off(c->P.NL) + 8);
    if (!c->peer_scratch) { close_peers(c); return fail(c, BOS_ERR_NOMEM, "peer scratch allocation failed"); }
    c->peers_open = true;
    return BOS_OK;
}

int bos_set_edge_shard(bos_ctx* c, int rank, int nranks) {
    if (!c || nranks < 1 || rank < 0 || rank >= nranks || nranks > 64) return BOS_ERR_INVALID;
    c->rank = rank; c->nranks = nranks;
    if (c->have_problem) compute_shard(c);
    return BOS_OK;
}

int bos_get_edge_shard(bos_ctx* c, int64_t* b_begin, int64_t* b_end, int64_t* o_begin, int64_t* o_end) {
    if (!c) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "shard before upload_problem");
    if (b_begin) *b_begin = c->shard.b_begin;
    if (b_end) *b_end = c->shard.b_end;
    if (o_begin) *o_begin = c->shard.o_begin;
    if (o_end) *o_end = c->shard.o_end;
    return BOS_OK;
}

}  // extern "C"

struct bos_host_pattern {
    HostPattern P;
};

extern "C" {

int bos_host_pattern_create(int NP, int NL, int fixed, int64_t Eb, const int32_t* b_pose, const int32_t* b_lm, int64_t Eo,
                            const int32_t* o_src, const int32_t* o_dst, bos_host_pattern** out) {
    if (!out) return BOS_ERR_INVALID;
    *out = nullptr;
    if ((Eb > 0 && (!b_pose || !b_lm)) || (Eo > 0 && (!o_src || !o_dst))) return BOS_ERR_INVALID;
    std::unique_ptr<bos_host_pattern> h(new bos_host_pattern());
    if (build_pattern(h->P, NP, NL, fixed, Eb, b_pose, b_lm, Eo, o_src, o_dst) != 0) return BOS_ERR_INVALID;
    *out = h.release();
    return BOS_OK;
}
int bos_host_pattern_destroy(bos_host_pattern* p) {
    delete p;
    return BOS_OK;
}
int bos_host_pattern_skyline(const bos_host_pattern* p, int32_t* panel_end, int32_t* n_panels, int32_t* rows_per_column, double* fill) {
    if (!p) return BOS_ERR_INVALID;
    DenseWork<double> w;
    skyline_symbolic<double>(p->P, w);
    if (n_panels) *n_panels = (int32_t)w.sky_panel_end.size();
    if (rows_per_column) *rows_per_column = w.sky_W;
    if (fill) *fill = w.sky_fill;
    if (panel_end) std::copy(w.sky_panel_end.begin(), w.sky_panel_end.end(), panel_end);
    return BOS_OK;
}
static uint64_t pattern_checksum(const HostPattern& P);
int bos_host_pattern_checksum(const bos_host_pattern* p, uint64_t* out) {
    if (!p || !out) return BOS_ERR_INVALID;
    *out = pattern_checksum(p->P);
    return BOS_OK;
}
int bos_pattern_checksum(bos_ctx* c, uint64_t* out) {
    if (!c || !out) return BOS_ERR_INVALID;
    NEED(c, c->have_problem, "pattern_checksum before upload_problem");
    *out = pattern_checksum(c->P);
    return BOS_OK;
}
int bos_set_device_setup(bos_ctx* c, int on) {
    if (!c) return BOS_ERR_INVALID;
    c->device_setup = on < 0 ? -1 : (on ? 1 : 0);
    return BOS_OK;
}
int bos_last_setup_ms(const bos_ctx* c, double* device_core_ms, double* host_ms) {
    if (!c) return BOS_ERR_INVALID;
    if (device_core_ms) *device_core_ms = c->device_setup_used ? c->setup_ms[0] : 0.0;
    if (host_ms) *host_ms = c->setup_ms[1];
    return BOS_OK;
}
int bos_device_resolve_ids(int device, int NP, const int32_t* pose_ids, int64_t Eb, const int32_t* b_pose_id, const int32_t* b_lm_id, int64_t Eo,
                           const int32_t* o_src_id, const int32_t* o_dst_id, int32_t* b_pose, int32_t* b_lm, int32_t* o_src, int32_t* o_dst,
                           int32_t* lm_ids, int32_t* NL_out) {
    if (!pose_ids || (Eb > 0 && (!b_pose_id || !b_lm_id || !b_pose || !b_lm)) || (Eo > 0 && (!o_src_id || !o_dst_id || !o_src || !o_dst))) return BOS_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return BOS_ERR_CUDA;
    std::string err;
    const int rc = device_resolve_ids(NP, pose_ids, Eb, b_pose_id, b_lm_id, Eo, o_src_id, o_dst_id, b_pose, b_lm, o_src, o_dst, lm_ids, NL_out, 0, err);
    return rc == 0 ? BOS_OK : (rc == 1 ? BOS_ERR_INVALID : BOS_ERR_CUDA);
}
static uint64_t pattern_checksum(const HostPattern& P) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* data, size_t bytes) {
        const unsigned char* b = static_cast<const unsigned char*>(data);
        for (size_t i = 0; i < bytes; i++) { h ^= b[i]; h *= 1099511628211ull; }
    };
    auto vec = [&](const auto& v) {
        const uint64_t n = v.size();
        mix(&n, sizeof(n));
        if (n) mix(v.data(), n * sizeof(v[0]));
    };
    const int scal[] = {P.NP, P.NL, P.fixed, P.Eb, P.Eo, P.N, P.pc_chunks, P.pc_cp, (int)P.pc_ok, (int)P.slots_identity, (int)P.has_shared_off};
    mix(scal, sizeof(scal));
    vec(P.b_pose); vec(P.b_lm); vec(P.b_perm); vec(P.b_slot); vec(P.slot_pose); vec(P.slot_lm);
    vec(P.pose_ptr); vec(P.lm_ptr); vec(P.lm_order); vec(P.lm_order_pose); vec(P.lm_order_lm);
    vec(P.o_src); vec(P.o_dst); vec(P.o_slot); vec(P.oe_ptr); vec(P.oe_edge); vec(P.oe_other); vec(P.o_shared);
    vec(P.off_lo); vec(P.off_hi); vec(P.pp_ptr); vec(P.pp_nbr); vec(P.pp_slot); vec(P.tri_ptr); vec(P.tri_edge);
    vec(P.pl_lm_id); vec(P.b_row); vec(P.ell_Loff); vec(P.ell_Lmap); vec(P.ell_Lpose);
    vec(P.pc_row_pose); vec(P.pc_goff); vec(P.pc_cl_ptr); vec(P.pc_cl_row); vec(P.pc_emap); vec(P.pc_nbr); vec(P.pc_nslot); vec(P.pc_ncnt);
    vec(P.lc_gptr); vec(P.lc_goff); vec(P.lc_emap); vec(P.sh_ptr); vec(P.sh_src); vec(P.sh_ell); vec(P.lc_row); vec(P.lc_k); vec(P.sh_first);
    vec(P.pc_loc); vec(P.tile_ptr); vec(P.tg_lm); vec(P.tg_eptr); vec(P.epose_ptr); vec(P.tg_edge); vec(P.touched);
    return h;
}
int bos_host_pattern_info(const bos_host_pattern* p, bos_pattern_info* out) {
    if (!p || !out) return BOS_ERR_INVALID;
    build_csc(const_cast<HostPattern&>(p->P));
    const HostPattern& P = p->P;
    out->n_hpl = (int64_t)P.slot_pose.size();
    out->n_hpp_off = (int64_t)P.off_lo.size();
    out->csc_n = P.N - 3;
    out->csc_nnz = (int64_t)P.csc_rowidx.size();
    out->N = P.N;
    out->vals_len = (int64_t)P.N + 6LL * P.NP + 3LL * P.NL + 9LL * (int64_t)P.off_lo.size() + 6LL * (int64_t)P.slot_pose.size();
    return BOS_OK;
}
int bos_host_pattern_get(const bos_host_pattern* p, int32_t* hpl_pose, int32_t* hpl_lm, int32_t* off_lo, int32_t* off_hi,
                         int64_t* b_slot, int64_t* o_slot, int32_t* csc_colptr, int32_t* csc_rowidx) {
    if (!p) return BOS_ERR_INVALID;
    build_csc(const_cast<HostPattern&>(p->P));
    const HostPattern& P = p->P;
    if (hpl_pose) std::copy(P.slot_pose.begin(), P.slot_pose.end(), hpl_pose);
    if (hpl_lm) std::copy(P.slot_lm.begin(), P.slot_lm.end(), hpl_lm);
    if (off_lo) std::copy(P.off_lo.begin(), P.off_lo.end(), off_lo);
    if (off_hi) std::copy(P.off_hi.begin(), P.off_hi.end(), off_hi);
    if (b_slot)
        for (int k = 0; k < P.Eb; k++) b_slot[P.b_perm[k]] = P.b_slot[k];
    if (o_slot)
        for (int e = 0; e < P.Eo; e++) o_slot[e] = P.o_slot[e];
    if (csc_colptr) std::copy(P.csc_colptr.begin(), P.csc_colptr.end(), csc_colptr);
    if (csc_rowidx) std::copy(P.csc_rowidx.begin(), P.csc_rowidx.end(), csc_rowidx);
    return BOS_OK;
}
int bos_host_edge_shard(int64_t Eb, int64_t Eo, int rank, int nranks, int64_t* out4) {
    if (!out4 || nranks < 1 || rank < 0 || rank >= nranks || Eb < 0 || Eo < 0) return BOS_ERR_INVALID;
    shard_ranges(Eb, Eo, rank, nranks, out4, nullptr);
    return BOS_OK;
}

}  // extern "C"

// ---- batched problems ---------------------------------------------------------------------------------------
struct bos_batch {
    bos_options opt;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevAlloc mem;
    BatchDev<double> dd;
    BatchDev<float> df;
    std::string err;
    bool f64() const { return opt.precision == BOS_PRECISION_F64; }
};

namespace {
template <typename S> BatchDev<S>& bdev(bos_batch* b);
template <> BatchDev<double>& bdev<double>(bos_batch* b) { return b->dd; }
template <> BatchDev<float>& bdev<float>(bos_batch* b) { return b->df; }

template <typename S>
int batch_create_impl(bos_batch* B, int nprob, int NP, int NL, int fixed, int Eb, const int32_t* b_pose, const int32_t* b_lm,
                      const double* b_z, const double* b_omega, int Eo, const int32_t* o_src, const int32_t* o_dst,
                      const double* o_z, const double* o_omega) {
    BatchDev<S>& d = bdev<S>(B);
    d.nprob = nprob; d.NP = NP; d.NL = NL; d.fixed = fixed; d.Eb = Eb; d.Eo = Eo;
    DevAlloc& m = B->mem;
    std::vector<int> bp(b_pose, b_pose + Eb), bl(b_lm, b_lm + Eb), os(o_src, o_src + Eo), od(o_dst, o_dst + Eo);
    std::vector<S> bom(Eb), bz = narrow<S>(b_z, (size_t)nprob * Eb), oz = narrow<S>(o_z, (size_t)nprob * Eo * 3), oom((size_t)6 * Eo);
    for (int e = 0; e < Eb; e++) bom[e] = b_omega ? (S)b_omega[e] : S(1);
    static const int up[6] = {0, 1, 2, 4, 5, 8};
    for (int e = 0; e < Eo; e++)
        for (int k = 0; k < 6; k++) oom[6 * (size_t)e + k] = (S)o_omega[9 * (size_t)e + up[k]];
    d.b_pose = m.upload(bp); d.b_lm = m.upload(bl); d.b_om = m.upload(bom); d.b_z = m.upload(bz);
    d.o_src = m.upload(os); d.o_dst = m.upload(od); d.o_om = m.upload(oom); d.o_z = m.upload(oz);
    d.pose = m.get<S>((size_t)nprob * NP * 4); d.lm = m.get<S>((size_t)nprob * std::max(NL, 1) * 2);
    d.chi2 = m.get<double>(2 * (size_t)nprob); d.delta_inf = m.get<double>((size_t)nprob); d.status = m.get<int>((size_t)nprob);
    if (!d.b_pose || !d.b_lm || !d.b_om || !d.b_z || !d.o_src || !d.o_dst || !d.o_om || !d.o_z || !d.pose || !d.lm || !d.chi2 ||
        !d.delta_inf || !d.status) {
        B->err = "device allocation failed";
        return BOS_ERR_NOMEM;
    }
    return BOS_OK;
}
template <typename S>
int batch_xfer(bos_batch* B, double* poses, double* lms, bool to_device) {
    BatchDev<S>& d = bdev<S>(B);
    const size_t np = (size_t)d.nprob * d.NP * 4, nl = (size_t)d.nprob * d.NL * 2;
    auto one = [&](S* dptr, double* h, size_t n) -> int {
        if (!h || n == 0) return 0;
        std::vector<S> v(n);
        if (to_device) {
            for (size_t i = 0; i < n; i++) v[i] = (S)h[i];
            if (cudaMemcpy(dptr, v.data(), n * sizeof(S), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        } else {
            if (cudaMemcpy(v.data(), dptr, n * sizeof(S), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
            for (size_t i = 0; i < n; i++) h[i] = (double)v[i];
        }
        return 0;
    };
    if (cudaStreamSynchronize(B->stream) != cudaSuccess) return BOS_ERR_CUDA;
    if (one(d.pose, poses, np) || one(d.lm, lms, nl)) { B->err = "state copy failed"; return BOS_ERR_CUDA; }
    return BOS_OK;
}
}  // namespace

extern "C" {

int bos_batch_create(const bos_options* opts, int nprob, int NP, int NL, int fixed, int Eb, const int32_t* b_pose, const int32_t* b_lm,
                     const double* b_z, const double* b_omega, int Eo, const int32_t* o_src, const int32_t* o_dst,
                     const double* o_z, const double* o_omega, bos_batch** out) {
    if (!out) return BOS_ERR_INVALID;
    *out = nullptr;
    if (nprob <= 0 || NP <= 0 || NL < 0 || Eb < 0 || Eo < 0 || fixed < 0 || fixed >= NP) return BOS_ERR_INVALID;
    if ((Eb > 0 && (!b_pose || !b_lm || !b_z)) || (Eo > 0 && (!o_src || !o_dst || !o_z || !o_omega))) return BOS_ERR_INVALID;
    for (int e = 0; e < Eb; e++)
        if (b_pose[e] < 0 || b_pose[e] >= NP || b_lm[e] < 0 || b_lm[e] >= NL) return BOS_ERR_INVALID;
    for (int e = 0; e < Eo; e++)
        if (o_src[e] < 0 || o_src[e] >= NP || o_dst[e] < 0 || o_dst[e] >= NP || o_src[e] == o_dst[e]) return BOS_ERR_INVALID;
    std::unique_ptr<bos_batch> B(new bos_batch());
    if (opts) B->opt = *opts; else bos_default_options(&B->opt);
    const size_t smem = batch_smem_bytes(NP, NL, B->f64() ? 8 : 4);
    if (smem > 200 * 1024) return BOS_ERR_INVALID;  // "mini-sized" problems only: the whole system lives in shared memory
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BOS_ERR_CUDA;
    if (cudaSetDevice(B->opt.device) != cudaSuccess) return BOS_ERR_CUDA;
    if (cudaStreamCreateWithFlags(&B->stream, cudaStreamNonBlocking) != cudaSuccess) return BOS_ERR_CUDA;
    if (cudaEventCreate(&B->ev0) != cudaSuccess || cudaEventCreate(&B->ev1) != cudaSuccess) return BOS_ERR_CUDA;
    int rc = B->f64() ? batch_create_impl<double>(B.get(), nprob, NP, NL, fixed, Eb, b_pose, b_lm, b_z, b_omega, Eo, o_src, o_dst, o_z, o_omega)
                      : batch_create_impl<float>(B.get(), nprob, NP, NL, fixed, Eb, b_pose, b_lm, b_z, b_omega, Eo, o_src, o_dst, o_z, o_omega);
    if (rc) return rc;
    *out = B.release();
    return BOS_OK;
}

int bos_batch_destroy(bos_batch* B) {
    if (!B) return BOS_OK;
    cudaSetDevice(B->opt.device);
    if (B->stream) cudaStreamSynchronize(B->stream);
    B->mem.release();
    if (B->ev0) cudaEventDestroy(B->ev0);
    if (B->ev1) cudaEventDestroy(B->ev1);
    if (B->stream) cudaStreamDestroy(B->stream);
    delete B;
    return BOS_OK;
}

const char* bos_batch_last_error(const bos_batch* B) { return B ? B->err.c_str() : "null batch"; }

int bos_batch_set_states(bos_batch* B, const double* poses, const double* lms) {
    if (!B) return BOS_ERR_INVALID;
    cudaSetDevice(B->opt.device);
    return B->f64() ? batch_xfer<double>(B, const_cast<double*>(poses), const_cast<double*>(lms), true)
                    : batch_xfer<float>(B, const_cast<double*>(poses), const_cast<double*>(lms), true);
}
int bos_batch_get_states(bos_batch* B, double* poses, double* lms) {
    if (!B) return BOS_ERR_INVALID;
    cudaSetDevice(B->opt.device);
    return B->f64() ? batch_xfer<double>(B, poses, lms, false) : batch_xfer<float>(B, poses, lms, false);
}

int bos_batch_step_device(bos_batch* B, int n_steps, float* elapsed_ms) {
    if (!B || n_steps < 0) return BOS_ERR_INVALID;
    cudaSetDevice(B->opt.device);
    cudaEventRecord(B->ev0, B->stream);
    for (int i = 0; i < n_steps; i++) {
        int rc = B->f64() ? launch_batch_step<double>(B->dd, B->opt.kernel_threshold, B->opt.damping, B->stream)
                          : launch_batch_step<float>(B->df, B->opt.kernel_threshold, B->opt.damping, B->stream);
        if (rc < 0) { B->err = "batch launch configuration failed"; return BOS_ERR_CUDA; }
    }
    cudaEventRecord(B->ev1, B->stream);
    cudaError_t e = cudaStreamSynchronize(B->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { B->err = cudaGetErrorString(e); return BOS_ERR_CUDA; }
    if (elapsed_ms) cudaEventElapsedTime(elapsed_ms, B->ev0, B->ev1);
    return BOS_OK;
}

int bos_batch_step(bos_batch* B, double* chi2, double* delta_inf, int32_t* status) {
    if (!B) return BOS_ERR_INVALID;
    int rc = bos_batch_step_device(B, 1, nullptr);
    if (rc) return rc;
    const int nprob = B->f64() ? B->dd.nprob : B->df.nprob;
    const double* dchi = B->f64() ? B->dd.chi2 : B->df.chi2;
    const double* ddel = B->f64() ? B->dd.delta_inf : B->df.delta_inf;
    const int* dst = B->f64() ? B->dd.status : B->df.status;
    if (chi2 && cudaMemcpy(chi2, dchi, 2 * (size_t)nprob * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return BOS_ERR_CUDA;
    if (delta_inf && cudaMemcpy(delta_inf, ddel, (size_t)nprob * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return BOS_ERR_CUDA;
    if (status && cudaMemcpy(status, dst, (size_t)nprob * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return BOS_ERR_CUDA;
    return BOS_OK;
}

}  // extern "C"
