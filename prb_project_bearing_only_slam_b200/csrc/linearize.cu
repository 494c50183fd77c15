// linearize.cu -- K1/K2/K3: per-edge errors + Jacobians and the block-sparse H, b assembly.
//
// Replaces the two per-edge loops and the damping of Solver::step (slam/solver.cpp:28-69) and
// error_and_jacobian x2 (slam/solver_jacobians.cpp:9-168).  The reference merges an N x N sparse
// temporary into H for every edge; here every edge adds straight into precomputed block slots.
//
//   K2+K3  k_pose_odometry_init   one thread per pose: damping * I (H += damping * I, solver.cpp:64-69) plus the
//          contributions of the odometry edges incident to that pose, gathered through a CSR list and written
//          with PLAIN stores (no atomics, no separate zero-init pass).  J_dst = -J_src entry for entry, so one
//          M = J_s^T Omega J_s and one 3-vector serve the source block, the destination block and the
//          off-diagonal block.  The same kernel resets the landmark blocks and b_lm.
//   K1     k_linearize_bearing    edge-parallel over the (pose, landmark)-sorted SoA edge buffer, FOUR consecutive
//          edges per thread (256-bit vector loads of the index / measurement / omega arrays):
//          - pose-landmark 3x2 blocks: owned by the edge -> stored SoA (6 planes), each thread writes its four
//            consecutive entries of a plane with one 256-bit store (no atomics, no zero-init);
//          - pose 3x3 diagonal block + b_pose: accumulated in registers over the thread's run of equal poses,
//            then a warp-segmented reduction over the lanes' runs, one RED per value per run head;
//          - landmark 2x2 diagonal block + b_lm: RED per edge (a landmark's edges are scattered over the buffer);
//          - chi2 / over-threshold counts: warp + block reduction, one RED per CTA.
//
// The fixed pose (gauge, solver.cpp:72-73) is handled by zeroing its Jacobian blocks at the source:
// its rows/cols then hold only the damping and a zero rhs, which is the same linear system as
// deleting them (dx_fixed = 0) without any special case downstream.
#include "bos_internal.h"
#include "bos_math.cuh"

#include <cstdlib>

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

// ---- K2 + K3 ---------------------------------------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(128, 8) k_pose_odometry_init(Dev<S> d, int o_begin, int o_end, S kernel_threshold, S damping) {
    __shared__ double red[2][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double chi_acc = 0.0;
    int over_acc = 0;
    if (i < d.NL) {
        S* hl = d.Hll + 3LL * i;
        hl[0] = damping; hl[1] = S(0); hl[2] = damping;
        d.b[3LL * d.NP + 2LL * i] = S(0);
        d.b[3LL * d.NP + 2LL * i + 1] = S(0);
    }
    if (i < d.NP) {
        S h[6] = {damping, S(0), S(0), damping, S(0), damping};
        S b[3] = {S(0), S(0), S(0)};
        const size_t Eo = (size_t)d.Eo;
        const PoseV<S> Xi = load_pose<S>(d.pose, i);
        const int q0 = __ldg(d.oe_ptr + i), q1 = __ldg(d.oe_ptr + i + 1);
        for (int q = q0; q < q1; q++) {
            const int code = __ldg(d.oe_edge + q);
            const int other = __ldg(d.oe_other + q);
            const int e = code >> 1, role = code & 1;
            if (e < o_begin || e >= o_end) continue;
            const PoseV<S> Xo = load_pose<S>(d.pose, other);
            const int s = role ? other : i, t = role ? i : other;
            const PoseV<S> Xs = role ? Xo : Xi, Xd = role ? Xi : Xo;
            S om[6];
#pragma unroll
            for (int k = 0; k < 6; k++) om[k] = __ldg(d.o_om + k * Eo + e);
            S err[3], u0, u1;
            odometry_terms<S>(Xs, Xd, __ldg(d.o_z + e), __ldg(d.o_z + Eo + e), __ldg(d.o_z + 2 * Eo + e), err, u0, u1);
            const S chi = odometry_chi<S>(om, err);
            S scale = S(1);
            const bool over = chi > kernel_threshold;
            if (over) scale = sqrt(kernel_threshold / chi);   // scales the ERROR only (slam/solver.cpp:54-58)
            S M[6], v[3];
            odometry_normal_terms<S>(Xs.c, Xs.s, u0, u1, om, err, M, v, scale);
            const bool fs = (s == d.fixed), ft = (t == d.fixed);
            if (role == 0) {
                chi_acc += (double)chi;
                over_acc += over ? 1 : 0;
                if (!fs) {
#pragma unroll
                    for (int k = 0; k < 6; k++) h[k] += M[k];
                    b[0] += v[0]; b[1] += v[1]; b[2] += v[2];
                }
                // H[lo][hi] += J_lo^T Omega J_hi = -M (M symmetric, so the orientation does not matter)
                S* ho = d.Hoff + 9LL * __ldg(d.o_slot + e);
                const S z = (fs || ft) ? S(0) : S(1);
                const S m9[9] = {-M[0] * z, -M[1] * z, -M[2] * z, -M[1] * z, -M[3] * z, -M[4] * z, -M[2] * z, -M[4] * z, -M[5] * z};
                if (d.o_shared[e]) {
#pragma unroll
                    for (int k = 0; k < 9; k++) red_add(ho + k, m9[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 9; k++) ho[k] = m9[k];
                }
            } else if (!ft) {
#pragma unroll
                for (int k = 0; k < 6; k++) h[k] += M[k];
                b[0] -= v[0]; b[1] -= v[1]; b[2] -= v[2];
            }
        }
        S* hp = d.Hpp + 6LL * i;
#pragma unroll
        for (int k = 0; k < 6; k++) hp[k] = h[k];
        S* bp = d.b + 3LL * i;
        bp[0] = b[0]; bp[1] = b[1]; bp[2] = b[2];
    }
    double c = warp_sum(chi_acc), o = warp_sum((double)over_acc);
    if (lane == 0) { red[0][warp] = c; red[1][warp] = o; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cs = 0, os = 0;
        for (int w = 0; w < 4; w++) { cs += red[0][w]; os += red[1][w]; }
        if (cs != 0.0) atomicAdd(d.stats + 1, cs);
        if (os != 0.0) atomicAdd(d.stats + 3, os);
    }
}

// ---- K1 ----------------------------------------------------------------------------------------------------------
constexpr int kLinThreads = 256;
constexpr int kEPT = kLinTile / kLinThreads;  // consecutive edges per thread (2)
static_assert(kLinThreads * kEPT == kLinTile && kEPT == 2, "one CTA covers exactly one tile, two edges per thread");
#ifndef BOS_LIN_MINBLOCKS
#define BOS_LIN_MINBLOCKS 3
#endif
constexpr int kNT = 14;  // per-edge terms staged in shared memory: 5 landmark-side + 9 pose-side

// One CTA per tile of kLinTile sorted edges.
//   phase 1: every thread linearizes two consecutive edges, stores their pose-landmark blocks (128-bit stores into
//            the SoA planes) and stages the 14 landmark-/pose-side products in shared memory;
//   phase 2: one thread per distinct landmark of the tile (host-precomputed grouping) and one thread per pose run of the
//            tile sum their edges from shared memory and issue ONE set of REDs each.
// The grouping metadata of phase 2 is fetched before phase 1 so its latency hides behind the arithmetic.
template <typename S, bool kIdentSlots>
__global__ void __launch_bounds__(kLinThreads, BOS_LIN_MINBLOCKS) k_linearize_bearing(Dev<S> d, int e_begin, int e_end, S kernel_threshold, int dbg) {
    __shared__ double red[2][kLinThreads / 32];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S (*st)[kLinTile] = reinterpret_cast<S (*)[kLinTile]>(smem_raw);   // [kNT][kLinTile]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int ta = e_begin + tile * kLinTile;                 // e_begin is a multiple of the tile size
    const int tb = (ta + kLinTile < e_end) ? ta + kLinTile : e_end;
    const int e0 = ta + tid * kEPT;
    const bool any = e0 < tb;
    // ---- loads: this thread's two edges, then the phase-2 metadata --------------------------------------------------
    int p2[kEPT] = {-1, -1}, l2[kEPT] = {0, 0};
    S z2[kEPT] = {S(0), S(0)}, om2[kEPT] = {S(0), S(0)};
    if (any) {  // the SoA arrays are padded to a multiple of 4 edges
        loadN<kEPT>(d.b_pose + e0, p2); loadN<kEPT>(d.b_lm + e0, l2);
        loadN<kEPT>(d.b_z + e0, z2); loadN<kEPT>(d.b_om + e0, om2);
    }
    const int gt = ta / kLinTile;
    const int g0 = __ldg(d.tile_ptr + gt), g1 = __ldg(d.tile_ptr + gt + 1);
    const int pfirst = __ldg(d.b_pose + ta), plast = __ldg(d.b_pose + tb - 1);
    int gl = 0, ga = 0, gb = 0, ra = 0, rb = 0;
    if (g0 + tid < g1) { gl = __ldg(d.tg_lm + g0 + tid); ga = __ldg(d.tg_eptr + g0 + tid); gb = __ldg(d.tg_eptr + g0 + tid + 1); }
    const int rt = kLinThreads - 1 - tid;   // pose runs are taken from the top of the CTA, landmark groups from the bottom
    if (pfirst + rt <= plast) { ra = __ldg(d.epose_ptr + pfirst + rt); rb = __ldg(d.epose_ptr + pfirst + rt + 1); }
    // ---- phase 1 -------------------------------------------------------------------------------------------------------
    double chi_acc = 0.0;
    int over_acc = 0;
    S hpl[6][kEPT];
    S tv[kNT][kEPT];
    PoseV<S> X = {S(0), S(0), S(1), S(0)};
    int xpose = -1;
#pragma unroll
    for (int j = 0; j < kEPT; j++) {
        const bool valid = any && (e0 + j < tb);
        S J[5] = {S(0), S(0), S(0), S(0), S(0)};
        S err = S(0), om = S(0);
        const int p = p2[j], l = l2[j];
        if (valid) {
            if (p != xpose) { X = load_pose<S>(d.pose, p); xpose = p; }
            S lx, ly;
            load_lm<S>(d.lm, l, lx, ly);
            om = om2[j];
            bearing_terms<S>(X, lx, ly, z2[j], err, J);
            // threshold robust kernel: scales the ERROR only (slam/solver.cpp:37-41)
            const S chi = err * om * err;
            chi_acc += (double)chi;
            if (chi > kernel_threshold) { err *= sqrt(kernel_threshold / chi); over_acc++; }
            if (p == d.fixed) { J[0] = J[1] = J[2] = S(0); }
        }
        const S w0 = J[0] * om, w1 = J[1] * om, w2 = J[2] * om;  // (J^T omega), pose part
        const S w3 = J[3] * om, w4 = J[4] * om;                  // landmark part
        hpl[0][j] = w0 * J[3]; hpl[1][j] = w0 * J[4];
        hpl[2][j] = w1 * J[3]; hpl[3][j] = w1 * J[4];
        hpl[4][j] = w2 * J[3]; hpl[5][j] = w2 * J[4];
        tv[0][j] = w3 * J[3]; tv[1][j] = w3 * J[4]; tv[2][j] = w4 * J[4];
        tv[3][j] = w3 * err; tv[4][j] = w4 * err;
        tv[5][j] = w0 * J[0]; tv[6][j] = w0 * J[1]; tv[7][j] = w0 * J[2];
        tv[8][j] = w1 * J[1]; tv[9][j] = w1 * J[2]; tv[10][j] = w2 * J[2];
        tv[11][j] = w0 * err; tv[12][j] = w1 * err; tv[13][j] = w2 * err;
        if (!kIdentSlots && valid) {
            const long long s = __ldg(d.b_slot + e0 + j);
#pragma unroll
            for (int k = 0; k < 6; k++) red_add(d.Hpl + (long long)k * d.hpl_ld + s, hpl[k][j]);
        }
    }
#pragma unroll
    for (int k = 0; k < kNT; k++) {   // two adjacent edges -> one 128-bit (64-bit for float) shared store, conflict-free
        typedef typename Vec2T<S>::type V2;
        V2 v; v.x = tv[k][0]; v.y = tv[k][1];
        *reinterpret_cast<V2*>(&st[k][tid * kEPT]) = v;
    }
    if (kIdentSlots && any && !(dbg & 4)) {
#pragma unroll
        for (int k = 0; k < 6; k++) storeN<kEPT>(d.Hpl + (long long)k * d.hpl_ld + e0, hpl[k]);
    }
    __syncthreads();
    // ---- phase 2: landmark groups -------------------------------------------------------------------------------------------
    if (!(dbg & 1)) {
        for (int g = g0 + tid; g < g1; g += kLinThreads) {
            if (g != g0 + tid) { gl = __ldg(d.tg_lm + g); ga = __ldg(d.tg_eptr + g); gb = __ldg(d.tg_eptr + g + 1); }
            S v0 = S(0), v1 = S(0), v2 = S(0), v3 = S(0), v4 = S(0);
            for (int q = ga; q < gb; q++) {
                const int le = __ldg(d.tg_edge + q);
                v0 += st[0][le]; v1 += st[1][le]; v2 += st[2][le]; v3 += st[3][le]; v4 += st[4][le];
            }
            S* hl = d.Hll + 3LL * gl;
            red_add(hl + 0, v0); red_add(hl + 1, v1); red_add(hl + 2, v2);
            S* bl = d.b + 3LL * d.NP + 2LL * gl;
            red_add(bl + 0, v3); red_add(bl + 1, v4);
        }
    }
    // ---- phase 2: pose runs (taken from the top of the CTA so they overlap the landmark groups of the low threads) --------
    if (!(dbg & 2)) {
        for (int p = pfirst + rt; p <= plast; p += kLinThreads) {
            if (p != pfirst + rt) { ra = __ldg(d.epose_ptr + p); rb = __ldg(d.epose_ptr + p + 1); }
            const int a = (ra > ta ? ra : ta) - ta, b = (rb < tb ? rb : tb) - ta;
            if (p == d.fixed || a >= b) continue;
            S v[9];
#pragma unroll
            for (int k = 0; k < 9; k++) v[k] = S(0);
            for (int q = a; q < b; q++) {
#pragma unroll
                for (int k = 0; k < 9; k++) v[k] += st[5 + k][q];
            }
            S* hp = d.Hpp + 6LL * p;
#pragma unroll
            for (int k = 0; k < 6; k++) red_add(hp + k, v[k]);
            S* bp = d.b + 3LL * p;
            red_add(bp + 0, v[6]); red_add(bp + 1, v[7]); red_add(bp + 2, v[8]);
        }
    }
    // ---- chi2 / over-threshold ------------------------------------------------------------------------------------------
    double c = warp_sum(chi_acc), o = warp_sum((double)over_acc);
    if (lane == 0) { red[0][warp] = c; red[1][warp] = o; }
    __syncthreads();
    if (tid == 0) {
        double cs = 0, os = 0;
        for (int w = 0; w < kLinThreads / 32; w++) { cs += red[0][w]; os += red[1][w]; }
        if (cs != 0.0) atomicAdd(d.stats + 0, cs);
        if (os != 0.0) atomicAdd(d.stats + 2, os);
    }
}

template <typename S>
int launch_linearize(const Dev<S>& d, const ShardRange& r, double kernel_threshold, double damping_here,
                     bool zero_hpl, bool zero_hoff, int sm_count, cudaStream_t st) {
    int launches = 0;
    cudaMemsetAsync(d.stats, 0, 8 * sizeof(double), st);
    if (zero_hoff && d.n_off > 0) cudaMemsetAsync(d.Hoff, 0, sizeof(S) * 9 * (size_t)d.n_off, st);
    if (zero_hpl && d.n_hpl > 0) cudaMemsetAsync(d.Hpl, 0, sizeof(S) * 6 * (size_t)d.hpl_ld, st);
    {
        const int n = d.NP > d.NL ? d.NP : d.NL;
        k_pose_odometry_init<S><<<(n + 127) / 128, 128, 0, st>>>(d, r.o_begin, r.o_end, (S)kernel_threshold, (S)damping_here);
        launches++;
    }
    const int nb = r.b_end - r.b_begin;
    if (nb > 0) {
        const int tiles = (nb + kLinTile - 1) / kLinTile;
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("BOS_LIN_DEBUG"); dbg = e ? atoi(e) : 0; }
        const size_t smem = sizeof(S) * kNT * kLinTile;
        static bool attr_done[2] = {false, false};
        bool& done = attr_done[sizeof(S) == 8 ? 0 : 1];
        if (!done) {
            cudaFuncSetAttribute(k_linearize_bearing<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_linearize_bearing<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            done = true;
        }
        if (d.b_slot == nullptr)
            k_linearize_bearing<S, true><<<tiles, kLinThreads, smem, st>>>(d, r.b_begin, r.b_end, (S)kernel_threshold, dbg);
        else
            k_linearize_bearing<S, false><<<tiles, kLinThreads, smem, st>>>(d, r.b_begin, r.b_end, (S)kernel_threshold, dbg);
        launches++;
    }
    (void)sm_count;
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
    __shared__ double red[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0.0;
    if (i < d.NP) {
        S dx = d.delta[3LL * i], dy = d.delta[3LL * i + 1], dt = d.delta[3LL * i + 2];
        PoseV<S> X = load_pose<S>(d.pose, i);
        S sd, cd;
        sincos(dt, &sd, &cd);
        S* o = d.pose + 4LL * i;
        o[0] = (cd * X.x + (-sd) * X.y) + dx;
        o[1] = (sd * X.x + cd * X.y) + dy;
        o[2] = cd * X.c + (-sd) * X.s;
        o[3] = sd * X.c + cd * X.s;
        m = fmax(fabs((double)dx), fmax(fabs((double)dy), fabs((double)dt)));
    } else if (i < d.NP + d.NL) {
        const int j = i - d.NP;
        S dx = d.delta[3LL * d.NP + 2LL * j], dy = d.delta[3LL * d.NP + 2LL * j + 1];
        d.lm[2LL * j] += dx;
        d.lm[2LL * j + 1] += dy;
        m = fmax(fabs((double)dx), fabs((double)dy));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(BOS_FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) m = fmax(m, red[w]);
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(d.stats + 4), (unsigned long long)__double_as_longlong(m));
    }
}
template <typename S>
int launch_update(const Dev<S>& d, cudaStream_t st) {
    int n = d.NP + d.NL;
    if (n > 0) k_update<S><<<(n + 255) / 256, 256, 0, st>>>(d);
    return 1;
}

template int launch_linearize<double>(const Dev<double>&, const ShardRange&, double, double, bool, bool, int, cudaStream_t);
template int launch_linearize<float>(const Dev<float>&, const ShardRange&, double, double, bool, bool, int, cudaStream_t);
template int launch_edge_terms<double>(const Dev<double>&, double*, double*, double*, double*, cudaStream_t);
template int launch_edge_terms<float>(const Dev<float>&, float*, float*, float*, float*, cudaStream_t);
template int launch_update<double>(const Dev<double>&, cudaStream_t);
template int launch_update<float>(const Dev<float>&, cudaStream_t);

}  // namespace bos
