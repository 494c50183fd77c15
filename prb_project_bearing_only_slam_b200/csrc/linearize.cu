// linearize.cu -- K1/K2/K3: per-edge errors + Jacobians and the block-sparse H, b assembly.
//
// Replaces the two per-edge loops and the damping of Solver::step (slam/solver.cpp:28-69) and
// error_and_jacobian x2 (slam/solver_jacobians.cpp:9-168).  The reference merges an N x N sparse
// temporary into H for every edge; here every edge adds straight into precomputed block slots.
//
//   K3     k_hb_init              every diagonal block starts at damping * I and b at 0 (H += damping * I, solver.cpp:64-69): one coalesced
//          store pass over 9 scalars per pose and 5 per landmark.
//   K2     k_linearize_odometry   one thread per odometry edge: error, J_s (J_dst = -J_src entry for entry, so one M = J_s^T Omega J_s and one
//          3-vector serve the source block, the destination block and the off-diagonal block), robust kernel; the off-diagonal block is
//          stored plainly (its slot belongs to the edge), the two diagonal blocks and b take REDs.  Independent of K1.
//   K1     k_linearize_bearing_persistent   persistent CTAs walk tiles of 512 (pose, landmark)-sorted edges fetched by TMA bulk
//          copies into a 2-stage shared-memory ring:
//          - pose-landmark 3x2 blocks: owned by the edge -> stored SoA (6 planes), 128-bit coalesced stores, no atomics;
//          - per edge FOUR numbers are staged in shared memory (a bearing residual is scalar, so every block the edge touches
//            is an outer product of sqrt(omega) * J with itself or with sqrt(omega) * e; the products are formed while summing);
//          - landmark 2x2 blocks + b_lm: one thread per distinct landmark of the tile (host-precomputed tile-local grouping)
//            sums its edges from shared memory and issues ONE set of REDs;
//          - pose 3x3 blocks + b_pose: one thread per pose run of the tile sums its edges and issues one set of REDs (round 1 stored
//            them plainly / into side slots and had a third kernel read them back: 25 us at 25 % occupancy; REDs cost less);
//          - chi2 / over-threshold counts: warp + block reduction, one atomic per CTA.
//
// The fixed pose (gauge, solver.cpp:72-73) is handled by zeroing its Jacobian blocks at the source:
// its rows/cols then hold only the damping and a zero rhs, which is the same linear system as
// deleting them (dx_fixed = 0) without any special case downstream.
#include "bos_internal.h"
#include "bos_math.cuh"
#include "bos_tma.cuh"

#include <cstdio>


namespace bos {

// ---- 4-wide vector access: 256-bit for double (LDG/STG.E.ENL2.256 on sm_100a), 128-bit for float -----------
__device__ __forceinline__ void load4(const double* p, double v[4]) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void load4(const float* p, float v[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const int* p, int v[4]) {
    int4 t = __ldg(reinterpret_cast<const int4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void store4(double* p, const double v[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void store4(float* p, const float v[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// 2-wide flavours (128-bit for double)
__device__ __forceinline__ void load2(const double* p, double v[2]) { double2 t = __ldg(reinterpret_cast<const double2*>(p)); v[0] = t.x; v[1] = t.y; }
__device__ __forceinline__ void load2(const float* p, float v[2]) { float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
__device__ __forceinline__ void load2(const int* p, int v[2]) { int2 t = __ldg(reinterpret_cast<const int2*>(p)); v[0] = t.x; v[1] = t.y; }
__device__ __forceinline__ void store2(double* p, const double v[2]) { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }
__device__ __forceinline__ void store2(float* p, const float v[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
template <int N, typename T> __device__ __forceinline__ void loadN(const T* p, T* v) {
    if constexpr (N == 4) load4(p, v); else if constexpr (N == 2) load2(p, v); else v[0] = __ldg(p);
}
template <int N, typename T> __device__ __forceinline__ void storeN(T* p, const T* v) {
    if constexpr (N == 4) store4(p, v); else if constexpr (N == 2) store2(p, v); else p[0] = v[0];
}

// ---- K3 -------------------------------------------------------------------------------------------------------------
template <typename S>
__device__ __forceinline__ void hb_init_item(const Dev<S>& d, S damping, int i) {
    if (i < d.NL) {
        S* hl = d.Hll + 3LL * i;
        hl[0] = damping; hl[1] = S(0); hl[2] = damping;
        d.b[3LL * d.NP + 2LL * i] = S(0);
        d.b[3LL * d.NP + 2LL * i + 1] = S(0);
    } else if (i - d.NL < d.n_cut) {      // poses whose bearing run is cut by a tile boundary: their parts arrive by REDs
        const int p = __ldg(d.cut_pose + (i - d.NL));
        S* hp = d.Hpp + 6LL * p;
#pragma unroll
        for (int k = 0; k < 6; k++) hp[k] = S(0);
        d.b[3LL * p] = S(0); d.b[3LL * p + 1] = S(0); d.b[3LL * p + 2] = S(0);
    }
}
template <typename S>
__global__ void __launch_bounds__(256) k_hb_init(Dev<S> d, S damping) { hb_init_item<S>(d, damping, blockIdx.x * blockDim.x + threadIdx.x); }

// ---- K2 -------------------------------------------------------------------------------------------------------------
#ifndef BOS_ODO_THREADS
#define BOS_ODO_THREADS 128
#endif
constexpr int kOdoThreads = BOS_ODO_THREADS;
template <typename S>
__global__ void __launch_bounds__(kOdoThreads) k_linearize_odometry(Dev<S> d, int o_begin, int o_end, int s_begin, int s_end, S kernel_threshold, S damping_init, int all_hoff) {
    __shared__ double red[2][kOdoThreads / 32];
    // programmatic dependent launch: the bearing kernel that follows may start its CTAs (edge prefetch, phase 1 of its first tiles: nothing
    // of that reads what this kernel writes) while this grid is still running; it waits (griddepcontrol.wait) before its first phase 2
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = o_begin + blockIdx.x * blockDim.x + threadIdx.x;
    // K3 rides in the same launch: the threads past the last edge initialise the landmark blocks and the cut poses (nothing here reads them)
    if (e >= o_end) hb_init_item<S>(d, damping_init, e - o_end);
    double chi_acc = 0.0;
    int over_acc = 0;
    bool return_skip = false;
    if (e < o_end) {
        const size_t Eo = (size_t)d.Eo;
        const int s = __ldg(d.o_src + e), t = __ldg(d.o_dst + e);
        const PoseV<S> Xs = load_pose<S>(d.pose, s), Xd = load_pose<S>(d.pose, t);
        const S ths = __ldg(d.theta + s), thd = __ldg(d.theta + t);
        S om[6], z[3];
#pragma unroll
        for (int k = 0; k < 6; k++) om[k] = __ldg(d.o_om + k * Eo + e);
#pragma unroll
        for (int k = 0; k < 3; k++) z[k] = __ldg(d.o_z + k * Eo + e);
        const int slot = __ldg(d.o_slot + e);
        const bool shared = d.o_shared[e] != 0;
        S err[3], u0, u1;
        odometry_terms<S>(Xs, Xd, ths, thd, z[0], z[1], z[2], err, u0, u1);
        const S chi = odometry_chi<S>(om, err);
        S scale = S(1);
        const bool over = chi > kernel_threshold;
        if (over) scale = sqrt(kernel_threshold / chi);   // scales the ERROR only (slam/solver.cpp:54-58)
        S M[6], v[3];
        odometry_normal_terms<S>(Xs.c, Xs.s, u0, u1, om, err, M, v, scale);
        if (d.irls && over) {                             // opt-in IRLS: the weight belongs to Omega, so it scales J^T Omega J as well
#pragma unroll
            for (int k = 0; k < 6; k++) M[k] *= scale;
        }
        if (e >= s_begin && e < s_end) { chi_acc = (double)chi; over_acc = over ? 1 : 0; }   // statistics: this rank's share of the edges only
        const bool fs = (s == d.fixed), ft = (t == d.fixed);      // gauge: the fixed pose's Jacobian block is zero
        // M and v go to a per-edge scratch (SoA, coalesced); the pose threads of the bearing kernel add them to the two diagonal blocks
#pragma unroll
        for (int k = 0; k < 6; k++) d.Mv[k * Eo + e] = M[k];
#pragma unroll
        for (int k = 0; k < 3; k++) d.Mv[(6 + k) * Eo + e] = v[k];
        // H[lo][hi] += J_lo^T Omega J_hi = -M (M symmetric, so the orientation does not matter)
        if (!all_hoff && (e < s_begin || e >= s_end)) return_skip = true;      // several ranks: the off-diagonal block comes from the rank whose share holds the edge
        S* ho = d.Hoff + 9LL * slot;
        const S zf = (fs || ft) ? S(0) : S(1);
        const S m9[9] = {-M[0] * zf, -M[1] * zf, -M[2] * zf, -M[1] * zf, -M[3] * zf, -M[4] * zf, -M[2] * zf, -M[4] * zf, -M[5] * zf};
        if (return_skip) {
        } else if (shared) {
#pragma unroll
            for (int k = 0; k < 9; k++) red_add(ho + k, m9[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 9; k++) ho[k] = m9[k];
        }
    }
    double c = warp_sum(chi_acc), o = warp_sum((double)over_acc);
    if (lane == 0) { red[0][warp] = c; red[1][warp] = o; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cs = 0, os = 0;
        for (int w = 0; w < kOdoThreads / 32; w++) { cs += red[0][w]; os += red[1][w]; }
        if (cs != 0.0) atomicAdd(d.stats_k2 + 1, cs);   // d.stats, or the scratch the peer barrier publishes to every replica (reduce_mode 4)
        if (os != 0.0) atomicAdd(d.stats_k2 + 3, os);
    }
}

// ---- K1 ----------------------------------------------------------------------------------------------------------
constexpr int kLinThreads = 256;
constexpr int kEPT = kLinTile / kLinThreads;  // consecutive edges per thread (2)
static_assert(kLinThreads * kEPT == kLinTile && kEPT == 2, "one CTA covers exactly one tile, two edges per thread");
#ifndef BOS_LIN_MINBLOCKS
#define BOS_LIN_MINBLOCKS 4
#endif
constexpr int kNT = 4;  // per-edge terms staged in shared memory: sqrt(omega) * (J_lm[0], J_lm[1], J_theta) and sqrt(omega) * e

// A CTA walks tiles blockIdx.x, +gridDim.x, ... and the tile's edge
// data (pose / landmark indices, measurement, omega, the tile-local landmark grouping) arrives by TMA bulk copies into a
// 2-stage shared-memory ring, fetched one tile ahead: the HBM latency of the edge stream is off the critical path, the
// grid is exactly the number of resident CTAs (kLinPersistCtas per SM).
constexpr int kLinStages = 2;
constexpr int kLinPersistCtas = BOS_LIN_MINBLOCKS;   // 4: 64 registers per thread, the whole register file and 220 KB of shared memory per SM
constexpr int kOeMax = 96;             // poses per tile whose odometry edge codes are staged with the tile (longer pose ranges read them from global memory)
template <typename S>
struct LinStage {
    int pose[kLinTile];
    int lm[kLinTile];
    S z[kLinTile];
    S om[kLinTile];
    unsigned short tge[kLinTile];          // tile-local edge index, grouped by landmark
    int glm[kLinTile + 8];                 // landmark of every group of the tile
    int pe[kLinTile + 8];                  // epose_ptr[p_lo ..]: bearing-edge ranges of the tile's poses
    unsigned short geptr[kLinTile + 8];    // tile-local range of every group in tge
    int hdr[4];                            // groups, first pose, last pose, p_lo
    int oec[2 * kOeMax];                   // oe2[p_lo ..]: the first two odometry edge codes of the tile's poses
};
template <typename S>
struct LinSmem {
    LinStage<S> stage[kLinStages];
    S st[kNT][kLinTile];
    double red[2][kLinThreads / 32];
    unsigned long long bar[kLinStages];
};

// Outputs that must be combined across ranks (landmark blocks, b, pose blocks, statistics).  Single rank / NCCL modes: the local buffer.
// kPeer (reduce_mode 4): the SAME store or RED goes to every rank's replica through its peer mapping (NVLink; the add is performed by the
// memory that owns the address, system scope), so when the last rank's kernel has finished every replica holds the combined H, b --
// the combine rides on the build, tile by tile, instead of following it as a collective.
template <bool kPeer, typename S>
__device__ __forceinline__ void out_red(const Dev<S>& d, S* p, S v) {
    if constexpr (!kPeer) red_add(p, v);
    else {
        const long long off = p - d.vals;
        for (int r = 0; r < d.npeer; r++) atomicAdd_system(d.pv[r] + off, v);
    }
}
template <bool kPeer, typename S>
__device__ __forceinline__ void out_store(const Dev<S>& d, S* p, S v) {
    if constexpr (!kPeer) *p = v;
    else {
        const long long off = p - d.vals;
        for (int r = 0; r < d.npeer; r++) d.pv[r][off] = v;
    }
}

template <typename S, bool kIdentSlots, bool kPeer = false>
__global__ void __launch_bounds__(kLinThreads, kLinPersistCtas) k_linearize_bearing_persistent(Dev<S> d, int e_begin, int e_end, S kernel_threshold, S damping) {
    extern __shared__ __align__(128) unsigned char lin_smem_raw[];
    LinSmem<S>& sm = *reinterpret_cast<LinSmem<S>*>(lin_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = (e_end - e_begin + kLinTile - 1) / kLinTile;
    constexpr unsigned kEdgeBytes = 2 * kLinTile * 4 + 2 * kLinTile * sizeof(S) + kLinTile * 2;
    auto issue = [&](int tile, int stg) {   // thread 0: everything the tile needs arrives by bulk copies on one mbarrier
        const size_t e = (size_t)e_begin + (size_t)tile * kLinTile;
        LinStage<S>& g = sm.stage[stg];
        const int4 m = __ldg(reinterpret_cast<const int4*>(d.tile_meta) + e / kLinTile);   // padded group offset, groups, first pose, last pose
        const int cnt = (m.y + 1 + 7) & ~7;
        int p_lo = m.z & ~3, pcnt = (m.w + 2 - p_lo + 3) & ~3;
        if (pcnt > kLinTile + 8) { p_lo = -1; pcnt = 0; }   // a long stretch of edge-free poses inside the tile: read epose_ptr directly
        g.hdr[0] = m.y; g.hdr[1] = m.z; g.hdr[2] = m.w; g.hdr[3] = p_lo;
        const bool oe = pcnt > 0 && pcnt <= kOeMax;
        mbar_expect_tx(&sm.bar[stg], kEdgeBytes + (unsigned)cnt * 6u + (unsigned)pcnt * (oe ? 12u : 4u));
        tma_bulk_load(g.pose, d.b_pose + e, kLinTile * 4, &sm.bar[stg]);
        tma_bulk_load(g.lm, d.b_lm + e, kLinTile * 4, &sm.bar[stg]);
        tma_bulk_load(g.z, d.b_z + e, kLinTile * sizeof(S), &sm.bar[stg]);
        tma_bulk_load(g.om, d.b_om + e, kLinTile * sizeof(S), &sm.bar[stg]);
        tma_bulk_load(g.tge, d.tg_edge + e, kLinTile * 2, &sm.bar[stg]);
        tma_bulk_load(g.glm, d.tgp_lm + m.x, (unsigned)cnt * 4u, &sm.bar[stg]);
        tma_bulk_load(g.geptr, d.tgp_eptr + m.x, (unsigned)cnt * 2u, &sm.bar[stg]);
        if (pcnt > 0) tma_bulk_load(g.pe, d.epose_ptr + p_lo, (unsigned)pcnt * 4u, &sm.bar[stg]);
        if (oe) tma_bulk_load(g.oec, d.oe2 + 2 * (size_t)p_lo, (unsigned)pcnt * 8u, &sm.bar[stg]);
    };
    if (tid == 0) {
        for (int s = 0; s < kLinStages; s++) mbar_init(&sm.bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        if ((int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
        if ((int)blockIdx.x + (int)gridDim.x < ntiles) issue(blockIdx.x + gridDim.x, 1);
    }
    double chi_acc = 0.0;
    int over_acc = 0;
    int k = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
        const int stg = k % kLinStages;
        const LinStage<S>& g = sm.stage[stg];
        const int ta = e_begin + tile * kLinTile;
        const int tb = (ta + kLinTile < e_end) ? ta + kLinTile : e_end;
        const int e0 = ta + tid * kEPT;
        const bool any = e0 < tb;
        mbar_wait(&sm.bar[stg], (unsigned)((k / kLinStages) & 1));
        // ---- phase 1 ---------------------------------------------------------------------------------------------------
        S hpl[6][kEPT];
        S tv[kNT][kEPT];
        PoseV<S> X = {S(0), S(0), S(1), S(0)};
        int xpose = -1;
#pragma unroll
        for (int j = 0; j < kEPT; j++) {
            const bool valid = any && (e0 + j < tb);
            S J[5] = {S(0), S(0), S(0), S(0), S(0)};
            S err = S(0), so = S(0);
            const int p = valid ? g.pose[tid * kEPT + j] : -1;
            if (valid) {
                const int l = g.lm[tid * kEPT + j];
                if (p != xpose) { X = load_pose<S>(d.pose, p); xpose = p; }
                S lx, ly;
                load_lm<S>(d.lm, l, lx, ly);
                const S om = g.om[tid * kEPT + j];
                bearing_terms<S>(X, lx, ly, g.z[tid * kEPT + j], err, J);
                // threshold robust kernel: scales the ERROR only (slam/solver.cpp:37-41)
                const S chi = err * om * err;
                chi_acc += (double)chi;
                so = (om == S(1)) ? S(1) : sqrt(om);
                if (chi > kernel_threshold) {
                    const S wgt = sqrt(kernel_threshold / chi);
                    over_acc++;
                    if (d.irls) so *= sqrt(wgt);   // opt-in IRLS: sqrt(w omega) on J and on e (b is the same, H is scaled by w)
                    else err *= wgt;
                }
            }
            const S j0 = so * J[3], j1 = so * J[4], jt = so * J[2];
            tv[0][j] = j0; tv[1][j] = j1; tv[2][j] = jt; tv[3][j] = so * err;
            const S fz = (valid && p == d.fixed) ? S(0) : S(1);   // gauge: the fixed pose's Jacobian block is zero
            hpl[0][j] = -(j0 * j0) * fz; hpl[1][j] = -(j0 * j1) * fz;
            hpl[2][j] = -(j1 * j0) * fz; hpl[3][j] = -(j1 * j1) * fz;
            hpl[4][j] = (jt * j0) * fz;  hpl[5][j] = (jt * j1) * fz;
            if (!kIdentSlots && valid) {
                const long long s = __ldg(d.b_slot + e0 + j);
#pragma unroll
                for (int q = 0; q < 6; q++) red_add(d.Hpl + (long long)q * d.hpl_ld + s, hpl[q][j]);
            }
        }
#pragma unroll
        for (int q = 0; q < kNT; q++) {
            typedef typename Vec2T<S>::type V2;
            V2 v; v.x = tv[q][0]; v.y = tv[q][1];
            *reinterpret_cast<V2*>(&sm.st[q][tid * kEPT]) = v;
        }
        if (kIdentSlots && any) {
#pragma unroll
            for (int q = 0; q < 6; q++) storeN<kEPT>(d.Hpl + (long long)q * d.hpl_ld + e0, hpl[q]);
        }
        __syncthreads();
        // everything below reads or accumulates into what the preceding kernel wrote (landmark blocks and cut poses initialised, the odometry
        // scratch Mv): wait for it -- once per CTA, a no-op when the kernel was not launched with programmatic stream serialization
        if (k == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        // ---- phase 2: ONE item list per tile -- the pose runs first (padded to whole warps, so that no warp mixes the two kinds of item),
        // then the landmark groups.  The runs are the long items (ten edges, then the owner's odometry parts): they start at once in the
        // first warps while the other warps sum the landmark groups beside them, instead of after them.
        const int ng = g.hdr[0], pfirst = g.hdr[1], plast = g.hdr[2], p_lo = g.hdr[3];
        constexpr int kRunSub = 2;   // lanes per pose run (they split the run's edges and the owner's odometry edges)
        const int nruns = plast - pfirst + 1, rpad = (nruns * kRunSub + 31) & ~31, nitems = rpad + ng;
        for (int it = tid; it < nitems; it += kLinThreads) {
            if (it >= rpad) {
                const int gi = it - rpad;
                S v0 = S(0), v1 = S(0), v2 = S(0), v3 = S(0), v4 = S(0);
                const int ga = g.geptr[gi], gb = g.geptr[gi + 1];
                for (int q = ga; q < gb; q++) {
                    const int le = g.tge[q];
                    const S j0 = sm.st[0][le], j1 = sm.st[1][le], ee = sm.st[3][le];
                    v0 += j0 * j0; v1 += j0 * j1; v2 += j1 * j1; v3 += j0 * ee; v4 += j1 * ee;
                }
                const int gl = g.glm[gi];
                S* hl = d.Hll + 3LL * gl;
                out_red<kPeer>(d, hl + 0, v0); out_red<kPeer>(d, hl + 1, v1); out_red<kPeer>(d, hl + 2, v2);
                S* bl = d.b + 3LL * d.NP + 2LL * gl;
                out_red<kPeer>(d, bl + 0, v3); out_red<kPeer>(d, bl + 1, v4);
            } else {
                const int run = it / kRunSub, sub = it % kRunSub;
                const bool have = run < nruns;
                const int p = have ? pfirst + run : plast;
                int ra, rb;
                if (p_lo >= 0) { ra = g.pe[p - p_lo]; rb = g.pe[p + 1 - p_lo]; }
                else { ra = __ldg(d.epose_ptr + p); rb = __ldg(d.epose_ptr + p + 1); }
                // The tile in which a pose's run STARTS owns the pose (edge-free poses included: their empty run starts somewhere too):
                // it adds the damping and the pose's odometry edges, so the block is final after ONE plain store.  Only a run cut by a
                // tile boundary (one pose per boundary, zeroed by k_hb_init) arrives in parts, by REDs.
                const bool owner = have && ((ra >= ta && ra < tb) || (tb == d.Eb && ra == d.Eb));
                const bool cut = rb > ra && (ra / kLinTile != (rb - 1) / kLinTile);
                const bool odo = owner && p != d.fixed;
                // The owner's odometry parts (K2's scratch) are fetched BEFORE the run is summed, into registers of their own, so the
                // latency of those loads hides behind the edge loop: lane `sub` takes the pose's sub-th odometry edge (its code comes
                // with the tile's staged data when the tile's pose range is short enough), lane 1 also the loop closures.
                S w[9];
#pragma unroll
                for (int q = 0; q < 9; q++) w[q] = S(0);
                int code = -1, q0 = 0, q1 = 0;
                const size_t Eo = (size_t)d.Eo;
                if (odo) {
                    const bool staged = p_lo >= 0 && ((plast + 2 - p_lo + 3) & ~3) <= kOeMax;
                    code = staged ? g.oec[2 * (p - p_lo) + sub] : __ldg(d.oe2 + 2 * (size_t)p + sub);
                    if (code >= 0) {       // (edge << 1) | role; J_dst = -J_src: M enters both diagonal blocks, v changes sign
                        const int e = code >> 1;
                        const S sg = (code & 1) ? S(-1) : S(1);
#pragma unroll
                        for (int q = 0; q < 6; q++) w[q] = __ldg(d.Mv + q * Eo + e);
#pragma unroll
                        for (int q = 0; q < 3; q++) w[6 + q] = sg * __ldg(d.Mv + (6 + q) * Eo + e);
                        if (sub == 1) { q0 = __ldg(d.oe_ptr + p) + 2; q1 = __ldg(d.oe_ptr + p + 1); }
                    }
                }
                S v[9];
#pragma unroll
                for (int q = 0; q < 9; q++) v[q] = S(0);
                if (owner && sub == 0) { v[0] = damping; v[3] = damping; v[5] = damping; }
                const int a = (ra > ta ? ra : ta) - ta, b = (rb < tb ? rb : tb) - ta;
                const bool live = have && (p != d.fixed) && a < b;
                if (live)
                    for (int q = a + sub; q < b; q += kRunSub) {   // J_pose = (-j0, -j1, jt)
                        const S j0 = sm.st[0][q], j1 = sm.st[1][q], jt = sm.st[2][q], ee = sm.st[3][q];
                        v[0] += j0 * j0; v[1] += j0 * j1; v[2] -= j0 * jt; v[3] += j1 * j1; v[4] -= j1 * jt; v[5] += jt * jt;
                        v[6] -= j0 * ee; v[7] -= j1 * ee; v[8] += jt * ee;
                    }
#pragma unroll
                for (int q = 0; q < 9; q++) v[q] += w[q];
                for (int q = q0; q < q1; q++) {                    // loop closures (lane 1 of an owner with two or more odometry edges)
                    const int cc = __ldg(d.oe_edge + q), e = cc >> 1;
                    const S sg = (cc & 1) ? S(-1) : S(1);
#pragma unroll
                    for (int k = 0; k < 6; k++) v[k] += __ldg(d.Mv + k * Eo + e);
#pragma unroll
                    for (int k = 0; k < 3; k++) v[6 + k] += sg * __ldg(d.Mv + (6 + k) * Eo + e);
                }
#pragma unroll
                for (int q = 0; q < 9; q++) v[q] += __shfl_xor_sync(BOS_FULL_MASK, v[q], 1);   // the branch is warp-uniform (rpad is a multiple of 32): all 32 lanes are here
                const bool any_live = have && (p != d.fixed) && a < b;   // the run has edges in this tile (whichever lane summed them)
                if (sub == 0 && (owner || any_live)) {
                    S* hp = d.Hpp + 6LL * p;
                    S* bp = d.b + 3LL * p;
                    if (cut) {
#pragma unroll
                        for (int q = 0; q < 6; q++) out_red<kPeer>(d, hp + q, v[q]);
                        out_red<kPeer>(d, bp, v[6]); out_red<kPeer>(d, bp + 1, v[7]); out_red<kPeer>(d, bp + 2, v[8]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 6; q++) out_store<kPeer>(d, hp + q, v[q]);
                        out_store<kPeer>(d, bp, v[6]); out_store<kPeer>(d, bp + 1, v[7]); out_store<kPeer>(d, bp + 2, v[8]);
                    }
                }
            }
        }
        __syncthreads();                       // everyone is done with this stage and with st
        if (tid == 0 && tile + kLinStages * (int)gridDim.x < ntiles) issue(tile + kLinStages * gridDim.x, stg);
    }
    // ---- chi2 / over-threshold: once per CTA ---------------------------------------------------------------------------------
    double c = warp_sum(chi_acc), o = warp_sum((double)over_acc);
    if (lane == 0) { sm.red[0][warp] = c; sm.red[1][warp] = o; }
    __syncthreads();
    if (tid == 0) {
        double cs = 0, os = 0;
        for (int w = 0; w < kLinThreads / 32; w++) { cs += sm.red[0][w]; os += sm.red[1][w]; }
        if constexpr (!kPeer) {
            if (cs != 0.0) atomicAdd(d.stats + 0, cs);
            if (os != 0.0) atomicAdd(d.stats + 2, os);
        } else {
            for (int r = 0; r < d.npeer; r++) {
                if (cs != 0.0) atomicAdd_system(d.pstats[r] + 0, cs);
                if (os != 0.0) atomicAdd_system(d.pstats[r] + 2, os);
            }
        }
    }
}

// Cross-GPU barrier of reduce_mode 4 (one warp, stream-ordered after the kernel whose remote writes it publishes: a kernel boundary makes
// them visible system-wide).  Lane q signals rank q's slot array, then waits for rank q's signal in its own; `publish` first adds the odometry
// kernel's share of the statistics to every replica.  A wait that exceeds ~4 s of clock raises the local error flag instead of hanging.
__global__ void k_peer_barrier(PeerBarrier pb) {
    const int q = threadIdx.x;
    if (pb.publish && q < pb.n) {
        const double c = pb.publish[1], o = pb.publish[3];
        if (c != 0.0) atomicAdd_system(pb.pstats[q] + 1, c);
        if (o != 0.0) atomicAdd_system(pb.pstats[q] + 3, o);
    }
    __threadfence_system();
    if (q < pb.n) {
        unsigned long long* dst = pb.slots[q] + pb.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(pb.epoch) : "memory");
        const unsigned long long* src = pb.slots[pb.rank] + q;
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
            if (v >= pb.epoch) break;
            if (clock64() - t0 > 8000000000LL) { *pb.error = 1ULL; break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}
int launch_peer_barrier(const PeerBarrier& pb, cudaStream_t st) {
    k_peer_barrier<<<1, 32, 0, st>>>(pb);
    return 1;
}

// ---- reduce_mode 5: pull-based combine over peer memory -----------------------------------------------------------------------
// scratch layout (S): [5 NL] landmark totals (Hll 3/landmark, then b_l 2/landmark) | [9 kMaxPeers] boundary-pose totals (Hpp 6, b 3) | [4] statistics (as S)
template <typename S>
__global__ void __launch_bounds__(256) k_peer_pull(Dev<S> d, PeerPull pp, S* __restrict__ T) {
    const long long nA = 6LL * pp.NP, nB = 3LL * pp.NP, nC = 5LL * pp.NL, nD = 9LL * (pp.n - 1), nE = 4;
    const long long oHpp = d.Hpp - d.vals, oHll = d.Hll - d.vals, oBl = 3LL * pp.NP;
    const long long gsz = (long long)gridDim.x * blockDim.x, gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    auto owner = [&](int p) { int q = 0; while (q + 1 < pp.n && p >= pp.own_p0[q + 1]) q++; return q; };
    // pose blocks / rhs: copied in place from the owner's replica (nobody reads this rank's copy of them); four independent NVLink loads in
    // flight per thread -- a remote load takes microseconds, the link is only busy when many of them overlap
    for (long long i0 = gid; i0 < nA + nB; i0 += 4 * gsz) {
        S v[4];
        long long off[4];
        bool take[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = i0 + u * gsz;
            take[u] = false; off[u] = 0; v[u] = S(0);
            if (i >= nA + nB) continue;
            const bool hp = i < nA;
            const long long e = hp ? i : i - nA;
            const int p = (int)(hp ? e / 6 : e / 3);
            const int q = owner(p);
            if (q == pp.rank || p == pp.bnd[q]) continue;
            off[u] = hp ? oHpp + e : e;
            take[u] = true;
            v[u] = __ldcg(d.pv[q] + off[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (take[u]) d.vals[off[u]] = v[u];
    }
    // landmark blocks and b_l: the sum of every rank's part, in rank order (identical on all ranks); all ranks' loads issued before the adds
    for (long long e = gid; e < nC; e += gsz) {
        const long long off = e < 3LL * pp.NL ? oHll + e : oBl + (e - 3LL * pp.NL);
        S part[kMaxPeers];
#pragma unroll
        for (int q = 0; q < kMaxPeers; q++) part[q] = q < pp.n ? __ldcg(d.pv[q] + off) : S(0);
        S s = S(0);
#pragma unroll
        for (int q = 0; q < kMaxPeers; q++) s += part[q];
        T[e] = s;
    }
    if (gid < nD) {                           // the pose whose run straddles the boundary between rank q and rank q + 1
        const int q = (int)(gid / 9), k = (int)(gid % 9), p = pp.bnd[q];
        if (p >= 0) {
            const long long off = k < 6 ? oHpp + 6LL * p + k : 3LL * p + (k - 6);
            T[nC + gid] = __ldcg(d.pv[q] + off) + __ldcg(d.pv[q + 1] + off);
        }
    } else if (gid < nD + nE) {               // chi2 / over-threshold counts of the ranks' shares
        const int k = (int)(gid - nD);
        double s = 0.0;
        for (int q = 0; q < pp.n; q++) s += __ldcg(d.pstats[q] + k);
        reinterpret_cast<double*>(T + peer_scratch_stats_off(pp.NL))[k] = s;   // 8-byte aligned slot behind the boundary totals
    }
}
template <typename S>
__global__ void __launch_bounds__(256) k_peer_commit(Dev<S> d, PeerPull pp, const S* __restrict__ T) {
    const long long nC = 5LL * pp.NL, nD = 9LL * (pp.n - 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nC + nD + 4; i += (long long)gridDim.x * blockDim.x) {
        if (i < nC) {
            if (i < 3LL * pp.NL) d.Hll[i] = T[i]; else d.b[3LL * pp.NP + (i - 3LL * pp.NL)] = T[i];
        } else if (i < nC + nD) {
            const long long e = i - nC;
            const int q = (int)(e / 9), k = (int)(e % 9), p = pp.bnd[q];
            if (p < 0) continue;
            if (k < 6) d.Hpp[6LL * p + k] = T[i]; else d.b[3LL * p + (k - 6)] = T[i];
        } else {
            const int k = (int)(i - nC - nD);
            d.stats[k] = reinterpret_cast<const double*>(T + peer_scratch_stats_off(pp.NL))[k];
        }
    }
}
template <typename S>
int launch_peer_pull(const Dev<S>& d, const PeerPull& pp, S* scratch, int sm_count, cudaStream_t st) {
    k_peer_pull<S><<<sm_count * 8, 256, 0, st>>>(d, pp, scratch);
    return 1;
}
template <typename S>
int launch_peer_commit(const Dev<S>& d, const PeerPull& pp, const S* scratch, cudaStream_t st) {
    const long long n = 5LL * pp.NL + 9LL * (pp.n - 1) + 4;
    k_peer_commit<S><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, pp, scratch);
    return 1;
}
template int launch_peer_pull<double>(const Dev<double>&, const PeerPull&, double*, int, cudaStream_t);
template int launch_peer_pull<float>(const Dev<float>&, const PeerPull&, float*, int, cudaStream_t);
template int launch_peer_commit<double>(const Dev<double>&, const PeerPull&, const double*, cudaStream_t);
template int launch_peer_commit<float>(const Dev<float>&, const PeerPull&, const float*, cudaStream_t);

// no bearing edges at all (a pure pose graph): nobody walks tiles, so the poses are finished here
template <typename S>
__global__ void __launch_bounds__(128) k_pose_finish_nobearing(Dev<S> d, S damping) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.NP) return;
    S v[9] = {damping, S(0), S(0), damping, S(0), damping, S(0), S(0), S(0)};
    if (p != d.fixed) {
        const size_t Eo = (size_t)d.Eo;
        for (int q = __ldg(d.oe_ptr + p); q < __ldg(d.oe_ptr + p + 1); q++) {
            const int code = __ldg(d.oe_edge + q), e = code >> 1;
            const S sg = (code & 1) ? S(-1) : S(1);
            for (int k = 0; k < 6; k++) v[k] += __ldg(d.Mv + k * Eo + e);
            for (int k = 0; k < 3; k++) v[6 + k] += sg * __ldg(d.Mv + (6 + k) * Eo + e);
        }
    }
    for (int k = 0; k < 6; k++) d.Hpp[6LL * p + k] = v[k];
    for (int k = 0; k < 3; k++) d.b[3LL * p + k] = v[6 + k];
}

template <typename S>
int launch_linearize(const Dev<S>& d, const ShardRange& r, double kernel_threshold, double damping, double damping_here,
                     bool zero_hpl, bool zero_hoff, int sm_count, cudaStream_t st, bool multi_rank, int rank, bool all_hoff, int phases) {
    // damping: what an owned pose block starts from; damping_here: what this rank adds to the landmark blocks (they are summed over ranks)
    // phases: bit 0 = initialisation + odometry kernel, bit 1 = bearing kernel (reduce_mode 4 puts a cross-GPU barrier between the two)
    int launches = 0;
    if (phases & 1) {
    cudaMemsetAsync(d.stats, 0, 16 * sizeof(double), st);   // statistics + the odometry kernel's scratch (tail of the value buffer)
    if (zero_hoff && d.n_off > 0) cudaMemsetAsync(d.Hoff, 0, sizeof(S) * 9 * (size_t)d.n_off, st);
    if (zero_hpl && d.n_hpl > 0) cudaMemsetAsync(d.Hpl, 0, sizeof(S) * 6 * (size_t)d.hpl_ld, st);
    // several ranks: a pose block is written by the rank that owns the pose only, the others contribute zeros to the combine
    if (multi_rank) cudaMemsetAsync(d.vals, 0, sizeof(S) * ((size_t)d.N + 6 * (size_t)d.NP), st);
    // every rank linearizes ALL odometry edges (0.2 M at synth-2M: the pose owner needs both of a pose's edges); statistics count the rank's share.
    // The same launch initialises the landmark blocks and the cut poses (K3).
    if (d.Eo > 0) {
        k_linearize_odometry<S><<<(d.Eo + d.NL + d.n_cut + kOdoThreads - 1) / kOdoThreads, kOdoThreads, 0, st>>>(d, 0, d.Eo, r.o_begin, r.o_end, (S)kernel_threshold, (S)damping_here, all_hoff ? 1 : 0);
        launches++;
    } else if (d.NL + d.n_cut > 0) { k_hb_init<S><<<(d.NL + d.n_cut + 255) / 256, 256, 0, st>>>(d, (S)damping_here); launches++; }
    }
    if (!(phases & 2)) return launches;
    const int nb = r.b_end - r.b_begin;
    if (nb > 0) {
        const int tiles = (nb + kLinTile - 1) / kLinTile;
        const size_t smem = sizeof(LinSmem<S>);
        ensure_dyn_smem((const void*)k_linearize_bearing_persistent<S, true>, smem);
        ensure_dyn_smem((const void*)k_linearize_bearing_persistent<S, false>, smem);
        int grid = sm_count * kLinPersistCtas;
        if (grid > tiles) grid = tiles;
        // launched with programmatic stream serialization: its CTAs may become resident while the odometry kernel is still running
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kLinThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const int eb = r.b_begin, ee = r.b_end;
        const S kt = (S)kernel_threshold, dm = (S)damping;
        if (d.peer_push && d.npeer > 0 && d.b_slot == nullptr) {
            ensure_dyn_smem((const void*)k_linearize_bearing_persistent<S, true, true>, smem);
            cudaLaunchKernelEx(&cfg, k_linearize_bearing_persistent<S, true, true>, d, eb, ee, kt, dm);
        } else if (d.b_slot == nullptr)
            cudaLaunchKernelEx(&cfg, k_linearize_bearing_persistent<S, true, false>, d, eb, ee, kt, dm);
        else
            cudaLaunchKernelEx(&cfg, k_linearize_bearing_persistent<S, false, false>, d, eb, ee, kt, dm);
        launches++;
    } else if (d.Eb == 0 && rank == 0 && d.NP > 0) {
        k_pose_finish_nobearing<S><<<(d.NP + 127) / 128, 128, 0, st>>>(d, (S)damping);
        launches++;
    }
    return launches;
}

// ---- per-edge error / Jacobian dump in the caller's edge order (parity tests) -------------------------
template <typename S>
__global__ void k_edge_terms_bearing(Dev<S> d, S* __restrict__ err_b, S* __restrict__ jac_b) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= d.Eb) return;
    int p = d.b_pose[e], l = d.b_lm[e];
    PoseV<S> X = load_pose<S>(d.pose, p);
    S lx, ly, err, J[5];
    load_lm<S>(d.lm, l, lx, ly);
    bearing_terms<S>(X, lx, ly, d.b_z[e], err, J);
    int o = d.b_perm[e];
    err_b[o] = err;
    for (int k = 0; k < 5; k++) jac_b[5LL * o + k] = J[k];
}
template <typename S>
__global__ void k_edge_terms_odometry(Dev<S> d, S* __restrict__ err_o, S* __restrict__ jac_o) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= d.Eo) return;
    const size_t Eo = (size_t)d.Eo;
    PoseV<S> Xs = load_pose<S>(d.pose, d.o_src[e]), Xd = load_pose<S>(d.pose, d.o_dst[e]);
    S err[3], u0, u1;
    odometry_terms<S>(Xs, Xd, d.o_z[e], d.o_z[Eo + e], d.o_z[2 * Eo + e], err, u0, u1);
    for (int k = 0; k < 3; k++) err_o[3LL * e + k] = err[k];
    S* J = jac_o + 18LL * e;
    const S c = Xs.c, s = Xs.s;
    S rows[3][3] = {{-c, -s, u0}, {s, -c, u1}, {S(0), S(0), S(-1)}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            J[i * 6 + j] = rows[i][j];
            J[i * 6 + 3 + j] = -rows[i][j];
        }
}
template <typename S>
int launch_edge_terms(const Dev<S>& d, S* err_b, S* jac_b, S* err_o, S* jac_o, cudaStream_t st) {
    if (d.Eb > 0) k_edge_terms_bearing<S><<<(d.Eb + 255) / 256, 256, 0, st>>>(d, err_b, jac_b);
    if (d.Eo > 0) k_edge_terms_odometry<S><<<(d.Eo + 255) / 256, 256, 0, st>>>(d, err_o, jac_o);
    return 2;
}

// ---- K7: State::apply_boxplus (framework/state.cpp:69-80, state.hpp:11-13) -----------------------------
// X <- v2t(dx) * X:  R' = R(dth) R, t' = R(dth) t + dt ; landmarks += dx.  Also max |dx| for the stats.
template <typename S>
__global__ void __launch_bounds__(256) k_update(Dev<S> d) {
    __shared__ double red[8], red2[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0.0, dig = 0.0;
    if (i < d.NP) {
        S dx = d.delta[3LL * i], dy = d.delta[3LL * i + 1], dt = d.delta[3LL * i + 2];
        S* o = d.pose + 4LL * i;
        const PoseV<S> X = {o[0], o[1], o[2], o[3]};   // plain (coherent) loads: this kernel rewrites the pose
        S sd, cd;
        sincos(dt, &sd, &cd);
        o[0] = (cd * X.x + (-sd) * X.y) + dx;
        o[1] = (sd * X.x + cd * X.y) + dy;
        o[2] = cd * X.c + (-sd) * X.s;
        o[3] = sd * X.c + cd * X.s;
        d.theta[i] = pose_theta<S>(PoseV<S>{o[0], o[1], o[2], o[3]});
        dig = (double)o[0] + (double)o[1] + (double)o[2] + (double)o[3];
        m = fmax(fabs((double)dx), fmax(fabs((double)dy), fabs((double)dt)));
    } else if (i < d.NP + d.NL) {
        const int j = i - d.NP;
        S dx = d.delta[3LL * d.NP + 2LL * j], dy = d.delta[3LL * d.NP + 2LL * j + 1];
        const S nx = d.lm[2LL * j] + dx, ny = d.lm[2LL * j + 1] + dy;
        d.lm[2LL * j] = nx;
        d.lm[2LL * j + 1] = ny;
        dig = (double)nx + (double)ny;
        m = fmax(fabs((double)dx), fabs((double)dy));
    }
    dig = warp_sum(dig);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(BOS_FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = m; red2[threadIdx.x >> 5] = dig; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { m = fmax(m, red[w]); dig += red2[w]; }
        atomicAdd(d.stats + 6, dig);
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(d.stats + 4), (unsigned long long)__double_as_longlong(m));
    }
}
// theta cache after the poses were replaced from the host (bos_set_state)
template <typename S>
__global__ void __launch_bounds__(256) k_pose_theta(Dev<S> d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.NP) d.theta[i] = pose_theta<S>(load_pose<S>(d.pose, i));
}
template <typename S>
int launch_pose_theta(const Dev<S>& d, cudaStream_t st) {
    if (d.NP > 0) k_pose_theta<S><<<(d.NP + 255) / 256, 256, 0, st>>>(d);
    return 1;
}
template <typename S>
int launch_update(const Dev<S>& d, cudaStream_t st) {
    int n = d.NP + d.NL;
    if (n > 0) k_update<S><<<(n + 255) / 256, 256, 0, st>>>(d);
    return 1;
}

template int launch_linearize<double>(const Dev<double>&, const ShardRange&, double, double, double, bool, bool, int, cudaStream_t, bool, int, bool, int);
template int launch_linearize<float>(const Dev<float>&, const ShardRange&, double, double, double, bool, bool, int, cudaStream_t, bool, int, bool, int);
template int launch_edge_terms<double>(const Dev<double>&, double*, double*, double*, double*, cudaStream_t);
template int launch_edge_terms<float>(const Dev<float>&, float*, float*, float*, float*, cudaStream_t);
template int launch_pose_theta<double>(const Dev<double>&, cudaStream_t);
template int launch_pose_theta<float>(const Dev<float>&, cudaStream_t);
template int launch_update<double>(const Dev<double>&, cudaStream_t);
template int launch_update<float>(const Dev<float>&, cudaStream_t);

}  // namespace bos
