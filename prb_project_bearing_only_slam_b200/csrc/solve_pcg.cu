// solve_pcg.cu -- K4/K5b/K6 for large problems: the 2x2 landmark blocks are Schur-complemented out
// IMPLICITLY and the reduced pose system  S dx_p = -(b_p - Hpl Hll^-1 b_l),
// S = Hpp - Hpl Hll^-1 Hlp, is solved by preconditioned conjugate gradients (3x3 block-Jacobi, or chunk-sized
// block-tridiagonal chain blocks + a Galerkin coarse space: see the fused variant below).
//
// Replaces SimplicialLDLT::factorize/solve on H_nofixed (slam/solver.hpp:72, slam/solver.cpp:77-94):
// eliminating the landmark blocks is exact, so the solution is the same dx up to the PCG tolerance.
// S is never formed (at 40 observations per landmark it would be ~1600 3x3 blocks per landmark).
// Classic variant (pcg_variant 1; a loop of small kernels): one application of S is two edge-parallel passes over the
// pose-landmark blocks:
//   t_l  = sum_k Hpl_k^T p_pose(k)      over the (landmark, pose)-ordered copy, run-reduced per landmark
//   y_p -= sum_k Hpl_k Hll^-1 t_lm(k)   over the (pose, landmark)-ordered blocks, run-reduced per pose
// both HBM-bound streams of 6 scalars per block; vectors and landmark blocks stay L2-resident.
#include "bos_internal.h"
#include "bos_math.cuh"
#include "bos_schur.cuh"
#include "bos_tma.cuh"

#include <cstdio>

namespace bos {

enum { SC_RZ0 = 0, SC_RZ1 = 1, SC_PAP = 2, SC_RZINIT = 3, SC_DONE = 4, SC_ITER = 5, SC_TOL2 = 6, SC_BAD = 7 };

__device__ __forceinline__ double block_sum_256(double v, double* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
    return s;  // valid in thread 0
}

// (lm, pose)-ordered copy of the pose-landmark blocks, refreshed once per GN iteration
template <typename S>
__global__ void __launch_bounds__(256) k_copy_hlp(int n_hpl, int ld, const S* __restrict__ Hpl, const int* __restrict__ order, S* __restrict__ Hlp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_hpl) return;
    const int s = __ldg(order + k);
#pragma unroll
    for (int c = 0; c < 6; c++) Hlp[(long long)c * ld + k] = Hpl[(long long)c * ld + s];
}

// one thread per pose: reduced rhs, diagonal block of S, its inverse, and the PCG start vectors
template <typename S>
__global__ void __launch_bounds__(256) k_pcg_pose_prep(Dev<S> d, PcgWork<S> w) {
    __shared__ double red[8];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double rz = 0.0;
    if (p < d.NP) {
        S g[3] = {-d.b[3LL * p], -d.b[3LL * p + 1], -d.b[3LL * p + 2]};
        S sd[6];
#pragma unroll
        for (int k = 0; k < 6; k++) sd[k] = d.Hpp[6LL * p + k];
        for (int s = d.pose_ptr[p]; s < d.pose_ptr[p + 1]; s++) {
            const S* B = d.Hpl + s;
            const int l = d.slot_lm[s];
            const S i00 = w.hllinv[3LL * l], i01 = w.hllinv[3LL * l + 1], i11 = w.hllinv[3LL * l + 2];
            const S u0 = w.ul[2LL * l], u1 = w.ul[2LL * l + 1];
            S b[6];
#pragma unroll
            for (int k = 0; k < 6; k++) b[k] = B[(long long)k * d.hpl_ld];
            S y[6];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                g[a] += b[2 * a] * u0 + b[2 * a + 1] * u1;
                y[2 * a] = b[2 * a] * i00 + b[2 * a + 1] * i01;
                y[2 * a + 1] = b[2 * a] * i01 + b[2 * a + 1] * i11;
            }
            sd[0] -= y[0] * b[0] + y[1] * b[1];
            sd[1] -= y[0] * b[2] + y[1] * b[3];
            sd[2] -= y[0] * b[4] + y[1] * b[5];
            sd[3] -= y[2] * b[2] + y[3] * b[3];
            sd[4] -= y[2] * b[4] + y[3] * b[5];
            sd[5] -= y[4] * b[4] + y[5] * b[5];
        }
        S mi[6];
        sym3_inverse<S>(sd, mi);
#pragma unroll
        for (int k = 0; k < 6; k++) w.minv[6LL * p + k] = mi[k];
        S z[3] = {mi[0] * g[0] + mi[1] * g[1] + mi[2] * g[2], mi[1] * g[0] + mi[3] * g[1] + mi[4] * g[2],
                  mi[2] * g[0] + mi[4] * g[1] + mi[5] * g[2]};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            w.x[3LL * p + a] = S(0);
            w.r[3LL * p + a] = g[a];
            w.p0[3LL * p + a] = z[a];
            rz += (double)g[a] * (double)z[a];
        }
    }
    double s = block_sum_256(rz, red);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(w.scal + SC_RZ0, s);
}

__global__ void k_pcg_begin(double* scal, double rtol) {
    scal[SC_RZINIT] = scal[SC_RZ0];
    scal[SC_TOL2] = rtol * rtol;
    scal[SC_DONE] = (scal[SC_RZ0] > 0.0) ? 0.0 : 1.0;
}

// y = Hpp p (diagonal blocks + pose-pose blocks); clears t_l and the scalars of this iteration
template <typename S>
__global__ void __launch_bounds__(256) k_pcg_y_init(Dev<S> d, PcgWork<S> w, int parity) {
    if (w.scal[SC_DONE] != 0.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { w.scal[SC_PAP] = 0.0; w.scal[parity ? SC_RZ0 : SC_RZ1] = 0.0; }
    for (long long k = i; k < 2LL * d.NL; k += (long long)gridDim.x * blockDim.x) w.tl[k] = S(0);
    if (i >= d.NP) return;
    const S* h = d.Hpp + 6LL * i;
    const S* pv = w.p0;
    const S x0 = pv[3LL * i], x1 = pv[3LL * i + 1], x2 = pv[3LL * i + 2];
    S y0 = h[0] * x0 + h[1] * x1 + h[2] * x2;
    S y1 = h[1] * x0 + h[3] * x1 + h[4] * x2;
    S y2 = h[2] * x0 + h[4] * x1 + h[5] * x2;
    for (int q = d.pp_ptr[i]; q < d.pp_ptr[i + 1]; q++) {
        const int nb = d.pp_nbr[q];
        const int sl = d.pp_slot[q];
        const S* B = d.Hoff + 9LL * (sl & 0x7fffffff);
        const S n0 = pv[3LL * nb], n1 = pv[3LL * nb + 1], n2 = pv[3LL * nb + 2];
        if (sl >= 0) {  // this pose is the row side of the stored block
            y0 += B[0] * n0 + B[1] * n1 + B[2] * n2;
            y1 += B[3] * n0 + B[4] * n1 + B[5] * n2;
            y2 += B[6] * n0 + B[7] * n1 + B[8] * n2;
        } else {        // column side: transpose
            y0 += B[0] * n0 + B[3] * n1 + B[6] * n2;
            y1 += B[1] * n0 + B[4] * n1 + B[7] * n2;
            y2 += B[2] * n0 + B[5] * n1 + B[8] * n2;
        }
    }
    w.y[3LL * i] = y0; w.y[3LL * i + 1] = y1; w.y[3LL * i + 2] = y2;
}

// y_p -= sum_k Hpl_k (Hll^-1 t_l), edge-parallel over the (pose, lm)-ordered blocks
template <typename S>
__global__ void __launch_bounds__(256) k_pcg_pose_scatter(Dev<S> d, PcgWork<S> w) {
    if (w.scal[SC_DONE] != 0.0) return;
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < d.n_hpl;
    int p = -1 - lane;
    S v[3] = {S(0), S(0), S(0)};
    if (valid) {
        p = __ldg(d.slot_pose + s);
        const int l = __ldg(d.slot_lm + s);
        const S t0 = w.tl[2LL * l], t1 = w.tl[2LL * l + 1];
        const S i00 = w.hllinv[3LL * l], i01 = w.hllinv[3LL * l + 1], i11 = w.hllinv[3LL * l + 2];
        const S u0 = i00 * t0 + i01 * t1, u1 = i01 * t0 + i11 * t1;
        const S* B = d.Hpl + s;
        const long long ld = d.hpl_ld;
        v[0] = -(B[0] * u0 + B[ld] * u1);
        v[1] = -(B[2 * ld] * u0 + B[3 * ld] * u1);
        v[2] = -(B[4 * ld] * u0 + B[5 * ld] * u1);
    }
    bool head;
    warp_run_reduce<S, 3>(v, p, lane, head);
    if (head && valid) {
        red_add(w.y + 3LL * p, v[0]);
        red_add(w.y + 3LL * p + 1, v[1]);
        red_add(w.y + 3LL * p + 2, v[2]);
    }
}

template <typename S>
__global__ void __launch_bounds__(256) k_pcg_dot(int n, PcgWork<S> w) {
    __shared__ double red[8];
    if (w.scal[SC_DONE] != 0.0) return;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)w.p0[i] * (double)w.y[i];
    double s = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(w.scal + SC_PAP, s);
}

// x += alpha p ; r -= alpha y ; z = M^-1 r ; rz_new += r.z
template <typename S>
__global__ void __launch_bounds__(256) k_pcg_update(Dev<S> d, PcgWork<S> w, int parity) {
    __shared__ double red[8];
    if (w.scal[SC_DONE] != 0.0) return;
    const double rz = w.scal[parity ? SC_RZ1 : SC_RZ0], pap = w.scal[SC_PAP];
    const S alpha = (pap > 0.0) ? (S)(rz / pap) : S(0);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (p < d.NP) {
        S r[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            w.x[3LL * p + a] += alpha * w.p0[3LL * p + a];
            r[a] = w.r[3LL * p + a] - alpha * w.y[3LL * p + a];
            w.r[3LL * p + a] = r[a];
        }
        const S* mi = w.minv + 6LL * p;
        S z[3] = {mi[0] * r[0] + mi[1] * r[1] + mi[2] * r[2], mi[1] * r[0] + mi[3] * r[1] + mi[4] * r[2],
                  mi[2] * r[0] + mi[4] * r[1] + mi[5] * r[2]};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            w.z[3LL * p + a] = z[a];
            acc += (double)r[a] * (double)z[a];
        }
    }
    double s = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(w.scal + (parity ? SC_RZ0 : SC_RZ1), s);
}

// p = z + beta p ; convergence test ; iteration count
template <typename S>
__global__ void __launch_bounds__(256) k_pcg_dir(int n, PcgWork<S> w, int parity) {
    if (w.scal[SC_DONE] != 0.0) return;
    const double rz = w.scal[parity ? SC_RZ1 : SC_RZ0], rz_new = w.scal[parity ? SC_RZ0 : SC_RZ1];
    const double pap = w.scal[SC_PAP];
    const S beta = (S)(rz_new / rz);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w.p0[i] = w.z[i] + beta * w.p0[i];
    __syncthreads();
    if (i == 0) {
        // every block has read DONE before block 0 can set it only if it is set after the grid-wide reads;
        // blocks that start late would skip their update, so the flag is written to a staging slot that
        // the NEXT kernel promotes (k_pcg_y_init reads SC_DONE; promotion happens in k_pcg_promote).
        w.scal[SC_ITER] += 1.0;
        bool stop = !(rz_new > w.scal[SC_TOL2] * w.scal[SC_RZINIT]);
        if (!(pap > 0.0)) { stop = true; w.scal[SC_BAD] = 1.0; }
        w.scal[8] = stop ? 1.0 : 0.0;
    }
}
__global__ void k_pcg_promote(double* scal) {
    if (scal[8] != 0.0) scal[SC_DONE] = 1.0;
}


// =================================================================================================================
// Fused variant (pcg_variant 0, default): the whole PCG solve is ONE persistent cooperative kernel (one 1024-thread CTA per
// SM, 2 grid barriers per CG iteration -- 3 with the coarse space --, no host round trips, no value atomics).
//
// Operator.  A bearing edge's 3x2 block is rank one, Hpl_k = Jp_k^T omega Jl_k, and its pose Jacobian is determined by
// the landmark Jacobian and the landmark position: Jp_k = (-j0, -j1, j0 ly - j1 lx) for Jl_k = (j0, j1)
// (slam/solver_jacobians.cpp:51-89: the pose translation columns are minus the landmark columns, the theta column is
// Jl . (ly, -lx)).  So S z = Hpp z - sum_k Jp_k^T (jh_k . u_l(k)),  u_l = Hll^-1 sum_k jh_k (Jp_k . z_pose(k)),
// jh = sqrt(omega) Jl, needs TWO scalars per edge instead of the six of the block.
//   landmark-major pass: sliced-ELL layout of pose indices (4 B per edge), 8 lanes per row, jh re-derived from the state;
//   pose-major pass:     one lane per pose; the lane holds its pose, each slot gathers one 32-byte landmark record
//                        (u_l and the landmark position) and RE-DERIVES jh from the state: 4 B per edge.
// The pose vectors p, s, x, r and the off-diagonal product live in SHARED MEMORY for the whole solve (a pose is owned by
// one lane of one CTA for all iterations); only z crosses CTAs.  Per CG iteration ~110 MB are touched at 2 M edges,
// all of it L2-resident.
//
// Recurrences: Chronopoulos-Gear CG (one reduction point per iteration): z = M^-1 r, w = S z, gamma = r.z, delta = z.w,
//   beta = gamma/gamma_old, alpha = gamma / (delta - beta gamma / alpha_old), p = z + beta p, s = w + beta s,
//   x += alpha p, r -= alpha s.   delta is assembled WITHOUT w:
//   delta = sum_i z_i.(Hpp_ii z_i) + sum_i z_i.(sum_nbr Hpp_ij z_j) - sum_l t_l.u_l
//   phase L: (a) owned poses: off-diagonal pose-pose products yoff_i = sum_nbr Hpp_ij z_j; (b) landmark rows:
//            t_l over the row's lanes, u_l = Hll^-1 t_l stored per landmark.
//   phase P: owned poses: w_i = Hpp_ii z_i + yoff_i - sum_k Jp_k^T (jh_k.u_l(k)) is complete locally, so the vector
//            updates, z' = M^-1 r and the next gamma / delta parts follow in the same thread.
enum { FS_GAMMA0 = 16, FS_DELTA0 = 19 };
constexpr int kPcgRows = (2048 + kPcgThreads - 1) / kPcgThreads;   // chunk rows per thread (a chunk has at most 2048 poses)
constexpr int kPcgSmemBudget = 227 * 1024 - 6144;   // dynamic shared memory available to the persistent kernel (5.7 KB are static)

// -DBOS_PCG_TIMING: thread 0 of a few CTAs prints clock64 deltas per phase (diagnostic builds only)
#ifdef BOS_PCG_TIMING
#define PCG_T(k) do { if (threadIdx.x == 0) { long long now__ = clock64(); tacc[k] += now__ - tlast; tlast = now__; } } while (0)
#else
#define PCG_T(k) do { } while (0)
#endif

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs of the (cooperatively launched, hence co-resident) grid; counter only grows, zeroed before the launch
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned nblocks, unsigned& epoch) {
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
        const unsigned target = epoch * nblocks;
#ifdef BOS_PCG_BARRIER_FENCE
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) { }
        __threadfence();
#else
        // release (cumulative over the CTA's writes ordered by the bar.sync above) / acquire pair, no separate fences
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        while (ld_acquire_u32(counter) < target) { }
#endif
    }
    __syncthreads();
}

// split form: arrive early (after the CTA's contribution is written by thread 0), wait later; same counter protocol
__device__ __forceinline__ void grid_arrive(unsigned* counter, unsigned& epoch) {
    epoch++;
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void grid_wait(unsigned* counter, unsigned nblocks, unsigned epoch) {
    if (threadIdx.x == 0) {
        const unsigned target = epoch * nblocks;
        while (ld_acquire_u32(counter) < target) { }
    }
    __syncthreads();
}

__device__ __forceinline__ double block_sum_pcg(double v, double* red) {   // result valid in thread 0
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    if (threadIdx.x == 0)
        for (int k = 0; k < kPcgThreads / 32; k++) s += red[k];
    return s;
}

// 4-padded records: one 32-byte (FP64) / 16-byte (FP32) sector per pose or landmark, read and written through L2 (.cg)
__device__ __forceinline__ void ld4cg(const double* p, double& a, double& b, double& c, double& d) {
    const double2 u = __ldcg(reinterpret_cast<const double2*>(p)), v = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    a = u.x; b = u.y; c = v.x; d = v.y;
}
__device__ __forceinline__ void ld4cg(const float* p, float& a, float& b, float& c, float& d) {
    const float4 u = __ldcg(reinterpret_cast<const float4*>(p));
    a = u.x; b = u.y; c = u.z; d = u.w;
}
__device__ __forceinline__ void st4cg(double* p, double a, double b, double c) {
    __stcg(reinterpret_cast<double2*>(p), make_double2(a, b));
    __stcg(reinterpret_cast<double2*>(p) + 1, make_double2(c, 0.0));
}
__device__ __forceinline__ void st4cg(float* p, float a, float b, float c) { __stcg(reinterpret_cast<float4*>(p), make_float4(a, b, c, 0.f)); }
__device__ __forceinline__ void st2cg(double* p, double a, double b) { __stcg(reinterpret_cast<double2*>(p), make_double2(a, b)); }
__device__ __forceinline__ void st2cg(float* p, float a, float b) { __stcg(reinterpret_cast<float2*>(p), make_float2(a, b)); }

// ---- once per GN iteration --------------------------------------------------------------------------------------------
// per compact landmark row: Hll^-1 and the record {Hll^-1 b_l, lx, ly}
template <typename S>
__global__ void __launch_bounds__(256) k_ell_fill(Dev<S> d, PcgWork<S> w) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < d.n_q) {   // per (chunk, landmark): what the persistent kernel needs of the landmark, in q order (coalesced per chunk)
        const int L = __ldg(d.pl_lm_id + __ldg(d.pc_cl_row + k));
        const size_t nq = (size_t)d.n_q;
        w.qstat[k] = w.hllinv[3LL * L]; w.qstat[nq + k] = w.hllinv[3LL * L + 1]; w.qstat[2 * nq + k] = w.hllinv[3LL * L + 2];
        w.qstat[3 * nq + k] = d.lm[2LL * L]; w.qstat[4 * nq + k] = d.lm[2LL * L + 1];
    }
    if (k < d.n_clm) {
        const int L = __ldg(d.pl_lm_id + k);
        w.hllinv_c[3LL * k] = w.hllinv[3LL * L]; w.hllinv_c[3LL * k + 1] = w.hllinv[3LL * L + 1]; w.hllinv_c[3LL * k + 2] = w.hllinv[3LL * L + 2];
        w.ul4[4LL * k] = w.ul[2LL * L]; w.ul4[4LL * k + 1] = w.ul[2LL * L + 1]; w.ul4[4LL * k + 2] = d.lm[2LL * L]; w.ul4[4LL * k + 3] = d.lm[2LL * L + 1];
    }
}

// one thread per chunk row (= pose): reduced rhs g = -(b_p - Hpl Hll^-1 b_l), the block-Jacobi preconditioner
// M_i = Hpp_ii - sum_k Hpl_k Hll^-1 Hpl_k^T (per EDGE: exact unless a (pose, landmark) pair is observed twice), the start
// vectors and the first gamma / delta parts
template <typename S>
__global__ void __launch_bounds__(256) k_pcg_fused_prep(Dev<S> d, PcgWork<S> w) {
    __shared__ double red[8];
    const long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)d.pc_chunks * d.pc_cp;
    double gz = 0.0, zw = 0.0;
    const int i = (R < nrows) ? __ldg(d.pc_row_pose + R) : -1;
    if (i >= 0) {
        const int c = (int)(R / d.pc_cp), r = (int)(R % d.pc_cp);
        const int gidx = c * (d.pc_cp / 32) + r / 32;
        const PoseV<S> X = load_pose<S>(d.pose, i);
        S gg[3] = {-d.b[3LL * i], -d.b[3LL * i + 1], -d.b[3LL * i + 2]};
        S hp[6], sd[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { hp[k] = d.Hpp[6LL * i + k]; sd[k] = hp[k]; }
        const int off = __ldg(d.pc_goff + gidx), W = __ldg(d.pc_goff + gidx + 1) - off;
        const int cl0 = __ldg(d.pc_cl_ptr + c);
        if (i != d.fixed)
            for (int t = 0; t < W; t++) {
                const long long slot = ((long long)off + t) * 32 + lane;
                const unsigned loc = d.pc_loc[slot];
                if (loc == 0xffffu) continue;
                const int row = __ldg(d.pc_cl_row + cl0 + (int)loc);
                const S u0 = w.ul4[4LL * row], u1 = w.ul4[4LL * row + 1], lx = w.ul4[4LL * row + 2], ly = w.ul4[4LL * row + 3];
                const S i00 = w.hllinv_c[3LL * row], i01 = w.hllinv_c[3LL * row + 1], i11 = w.hllinv_c[3LL * row + 2];
                S j0, j1;
                bearing_jl_world<S>(X.x, X.y, lx, ly, j0, j1);
                const S so = w.omega_uniform ? (S)w.sqrt_omega : __ldg(w.Pw + slot);
                j0 *= so; j1 *= so;
                const S jp[3] = {-j0, -j1, j0 * ly - j1 * lx};
                const S m = j0 * u0 + j1 * u1;
                const S q = i00 * j0 * j0 + S(2) * i01 * j0 * j1 + i11 * j1 * j1;
                gg[0] += jp[0] * m; gg[1] += jp[1] * m; gg[2] += jp[2] * m;
                sd[0] -= q * jp[0] * jp[0]; sd[1] -= q * jp[0] * jp[1]; sd[2] -= q * jp[0] * jp[2];
                sd[3] -= q * jp[1] * jp[1]; sd[4] -= q * jp[1] * jp[2]; sd[5] -= q * jp[2] * jp[2];
            }
        if (w.precond != 1) {   // chain preconditioner: its diagonal block and the block that couples this row to the next chunk row
            const int inext = (r + 1 < d.pc_cp) ? __ldg(d.pc_row_pose + R + 1) : -1;
            S o[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
            if (inext >= 0)
                for (int q = __ldg(d.pp_ptr + i); q < __ldg(d.pp_ptr + i + 1); q++)
                    if (__ldg(d.pp_nbr + q) == inext) {
                        const S* Bo = d.Hoff + 9LL * (__ldg(d.pp_slot + q) & 0x7fffffff);
                        o[0] = Bo[0]; o[1] = Bo[1]; o[2] = Bo[2]; o[3] = Bo[4]; o[4] = Bo[5]; o[5] = Bo[8];
                    }
#pragma unroll
            for (int k = 0; k < 6; k++) { w.chD[(size_t)k * nrows + R] = sd[k]; w.chO[(size_t)k * nrows + R] = o[k]; }
        }
        S mi[6];
        sym3_inverse<S>(sd, mi);
#pragma unroll
        for (int k = 0; k < 6; k++) w.minv[6LL * i + k] = mi[k];
        const S z[3] = {mi[0] * gg[0] + mi[1] * gg[1] + mi[2] * gg[2], mi[1] * gg[0] + mi[3] * gg[1] + mi[4] * gg[2],
                        mi[2] * gg[0] + mi[4] * gg[1] + mi[5] * gg[2]};
        const S hz[3] = {hp[0] * z[0] + hp[1] * z[1] + hp[2] * z[2], hp[1] * z[0] + hp[3] * z[1] + hp[4] * z[2],
                         hp[2] * z[0] + hp[4] * z[1] + hp[5] * z[2]};
        const size_t np4 = 4 * (size_t)d.NP;
#pragma unroll
        for (int a = 0; a < 3; a++) w.rS[(size_t)a * nrows + R] = gg[a];
#pragma unroll
        for (int k = 0; k < 6; k++) { w.rowS[(size_t)k * nrows + R] = hp[k]; w.rowS[(size_t)(6 + k) * nrows + R] = mi[k]; }
#pragma unroll
        for (int n = 0; n < 2; n++) {
            const int nb = __ldg(d.pc_nbr + (size_t)n * nrows + R);
            S o[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
            if (nb >= 0) {   // blocks are symmetric by construction (-J_s^T Omega J_s), the orientation flag does not matter
                const S* Bo = d.Hoff + 9LL * (__ldg(d.pc_nslot + (size_t)n * nrows + R) & 0x7fffffff);
                o[0] = Bo[0]; o[1] = Bo[1]; o[2] = Bo[2]; o[3] = Bo[4]; o[4] = Bo[5]; o[5] = Bo[8];
            }
#pragma unroll
            for (int k = 0; k < 6; k++) w.rowS[(size_t)(12 + 6 * n + k) * nrows + R] = o[k];
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            w.z4[4LL * i + a] = (a < 3 && w.precond == 1) ? z[a] : S(0);   // chain: the persistent kernel applies M^-1 to g itself
            w.z4[np4 + 4LL * i + a] = S(0);
        }
        if (w.precond == 1) {
#pragma unroll
            for (int a = 0; a < 3; a++) {
                gz += (double)gg[a] * (double)z[a];
                zw += (double)z[a] * (double)hz[a];
            }
        }
    } else if (R < nrows && w.precond != 1) {   // padding row: identity block, no coupling
#pragma unroll
        for (int k = 0; k < 6; k++) { w.chD[(size_t)k * nrows + R] = (k == 0 || k == 3 || k == 5) ? S(1) : S(0); w.chO[(size_t)k * nrows + R] = S(0); }
    }
    double s1 = block_sum_256(gz, red);
    __syncthreads();
    double s2 = block_sum_256(zw, red);
    if (threadIdx.x == 0) {
        if (s1 != 0.0) atomicAdd(w.scal + FS_GAMMA0, s1);
        if (s2 != 0.0) atomicAdd(w.scal + FS_DELTA0, s2);
    }
}

// Landmark rows of the L layout, kEllLanesL lanes per row; the per-edge factors are re-derived from the state.
// MODE 0: t_l from vec4 (= z), u_l = Hll^-1 t_l stored, dacc -= t.u.   MODE 1: dx_l = Hll^-1 (-b_l - t_l) with vec4 = x.
template <typename S, int MODE>
__device__ __forceinline__ void pcg_landmark_rows(const Dev<S>& d, const PcgWork<S>& w, const S* vec4, int wg, int nwarps, double& dacc) {
    constexpr int RPG = 32 / kEllLanesL;
    const int lane = threadIdx.x & 31;
    const S so_u = (S)w.sqrt_omega;
    for (int g = wg; g < d.nLg; g += nwarps) {
        const int off = __ldg(d.ell_Loff + g), W = __ldg(d.ell_Loff + g + 1) - off;
        const int row = g * RPG + lane / kEllLanesL;
        const bool valid = row < d.n_clm;
        S lx = S(0), ly = S(0);
        if (valid) { lx = __ldg(w.ul4 + 4LL * row + 2); ly = __ldg(w.ul4 + 4LL * row + 3); }   // static half of the record
        S t0 = S(0), t1 = S(0);
        const long long s0 = (long long)off * 32 + lane;
        for (int tb = 0; tb < W; tb += 8) {
            int psv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) psv[k] = (tb + k < W) ? __ldg(d.ell_Lpose + s0 + (long long)(tb + k) * 32) : -1;   // all index loads first
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int ps = psv[k];
                if (ps < 0) continue;                              // padding, or an edge of the fixed pose (zero Jacobian block)
                S px, py;
                load_lm<S>(d.pose, 2 * ps, px, py);                 // the pose translation: first half of the (x, y, c, s) record
                S z0, z1, z2, zp;
                ld4cg(vec4 + 4LL * ps, z0, z1, z2, zp);
                S j0, j1;
                bearing_jl_world<S>(px, py, lx, ly, j0, j1);
                const S so = w.omega_uniform ? so_u : __ldg(w.Lw + s0 + (long long)(tb + k) * 32);
                j0 *= so; j1 *= so;
                const S sc = (j0 * ly - j1 * lx) * z2 - j0 * z0 - j1 * z1;   // Jp_k . z
                t0 += j0 * sc; t1 += j1 * sc;
            }
        }
#pragma unroll
        for (int o = 1; o < kEllLanesL; o <<= 1) {
            t0 += __shfl_xor_sync(BOS_FULL_MASK, t0, o);
            t1 += __shfl_xor_sync(BOS_FULL_MASK, t1, o);
        }
        if (valid && (lane % kEllLanesL) == 0) {
            const S i00 = __ldg(w.hllinv_c + 3LL * row), i01 = __ldg(w.hllinv_c + 3LL * row + 1), i11 = __ldg(w.hllinv_c + 3LL * row + 2);
            if (MODE == 0) {
                const S u0 = i00 * t0 + i01 * t1, u1 = i01 * t0 + i11 * t1;
                st2cg(w.ul4 + 4LL * row, u0, u1);
                dacc -= (double)t0 * (double)u0 + (double)t1 * (double)u1;
            } else {
                const int L = __ldg(d.pl_lm_id + row);
                const S r0 = -d.b[3LL * d.NP + 2LL * L] - t0, r1 = -d.b[3LL * d.NP + 2LL * L + 1] - t1;
                d.delta[3LL * d.NP + 2LL * L] = i00 * r0 + i01 * r1;
                d.delta[3LL * d.NP + 2LL * L + 1] = i01 * r0 + i11 * r1;
            }
        }
    }
}

// Chunk-local landmark pass of the operator: the chunk's partial t_l = sum_k jh_k (Jp_k . z_pose(k)) over ITS OWN edges of every landmark it
// sees, with the poses' z (rows 9-11 of the shared vectors) and positions (pxy) read from shared memory; one 16-byte record per (chunk,
// landmark) goes to global memory, the chunks that share a landmark sum each other's records after the grid barrier (pattern.cpp "LC").
// Round 1 gathered pose position + z (two 32-byte sectors per EDGE) from L2 here: 30 of the 100 us of a CG iteration.
template <typename S>
__device__ __forceinline__ void pcg_local_landmark_rows(const Dev<S>& d, const PcgWork<S>& w, int c, const S* zs, int cps, int Kp, bool chain,
                                                        const S* pxy, S* tp, int* next_group) {
    constexpr int RPG = 32 / kLcLanes;   // rows (landmarks) per group
    const int lane = threadIdx.x & 31;
    const int sub = lane % kLcLanes, rowl = lane / kLcLanes;
    const int g0 = __ldg(d.lc_gptr + c), ng = __ldg(d.lc_gptr + c + 1) - g0;
    const int cl0 = __ldg(d.pc_cl_ptr + c), ncl = __ldg(d.pc_cl_ptr + c + 1) - cl0;
    const S so_u = (S)w.sqrt_omega;
    const size_t nq = (size_t)d.n_q;
    // The groups are sorted by descending length; the warps take them from a shared counter (*next_group, zero on entry) as they become free:
    // longest-processing-time-first scheduling.  A static round robin left the warps that drew two long groups working twice as long as the rest.
    for (;;) {
        int gi = 0;
        if (lane == 0) gi = atomicAdd(next_group, 1);
        gi = __shfl_sync(BOS_FULL_MASK, gi, 0);
        if (gi >= ng) break;
        const int g = g0 + gi;
        const int off = __ldg(d.lc_goff + g), W = __ldg(d.lc_goff + g + 1) - off;
        const int k = (int)__ldg(d.lc_k + (size_t)g0 * 32 + (size_t)gi * RPG + rowl);   // the local landmark of this lane's row
        const bool valid = k != 0xffff && k < ncl;
        S lx = S(0), ly = S(0);
        if (valid) { lx = __ldg(w.qstat + 3 * nq + cl0 + k); ly = __ldg(w.qstat + 4 * nq + cl0 + k); }
        S t0 = S(0), t1 = S(0);
        const long long s0 = (long long)off * 32 + lane;
        for (int tb = 0; tb < W; tb += 8) {
            unsigned rv[8];
#pragma unroll
            for (int q = 0; q < 8; q++) rv[q] = (tb + q < W) ? (unsigned)__ldg(d.lc_row + s0 + (long long)(tb + q) * 32) : 0xffffu;   // all index loads first
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const unsigned r = rv[q];
                if (r == 0xffffu) continue;
                const int xv = chain ? (int)(r & 31u) * Kp + (int)(r >> 5) : (int)r;
                const S px = pxy[2 * r], py = pxy[2 * r + 1];
                const S z0 = zs[xv], z1 = zs[cps + xv], z2 = zs[2 * cps + xv];
                S j0, j1;
                bearing_jl_world<S>(px, py, lx, ly, j0, j1);
                const S so = w.omega_uniform ? so_u : __ldg(w.Cw + s0 + (long long)(tb + q) * 32);
                j0 *= so; j1 *= so;
                const S sc = (j0 * ly - j1 * lx) * z2 - j0 * z0 - j1 * z1;   // Jp_k . z
                t0 += j0 * sc; t1 += j1 * sc;
            }
        }
#pragma unroll
        for (int o = 1; o < kLcLanes; o <<= 1) { t0 += __shfl_xor_sync(BOS_FULL_MASK, t0, o); t1 += __shfl_xor_sync(BOS_FULL_MASK, t1, o); }
        if (valid && sub == 0) st2cg(tp + 2LL * (cl0 + k), t0, t1);
    }
}

// ---- chain preconditioner ----------------------------------------------------------------------------------------------
// Per chunk (= CTA of the persistent kernel) M is the block-tridiagonal matrix of the Schur diagonal blocks D_r and the
// pose-pose blocks O_r between consecutive chunk rows (the odometry chain; every such block is -J_s^T Omega J_s, symmetric).
// M = H_chain + blockdiag(bearing Schur diagonal + damping) is SPD.  It is solved exactly with two levels: the chunk's rows
// form groups of 32 = 31 interior rows + 1 separator row; the interiors are independent block-tridiagonal systems (one thread
// each, block LDL^T), the separators a block-tridiagonal Schur system over the groups.  Factors are stored as FP32 (the applied
// operator stays symmetric positive definite whatever the rounding: it is a congruence of blockdiag(T~^-1, S~^-1)), in the
// layout the solve reads from shared memory: row (group g, position k) at k * Kp + g with Kp odd.
// Layout of a chunk's factor block (floats), read from shared memory by the solve as 16-byte vectors without bank conflicts:
//   interior row (group g, position k): four float4 at F + ((k * 4 + c) * Kp + g) * 4, c = 0..3 holding
//     {L0..L3} {L4..L7} {L8, w00, w10, w11} {w20, w21, w22, 0}   (L row-major 3x3 with L_0 = 0, W W^T = Delta^-1)
//   separator j: the same 16 floats contiguous at Fs + 16 * j, Fs = F + 16 * cps
//   coupling blocks of group g: three float4 at Fc + (c * Kp + g) * 4, Fc = Fs + 16 * Kp:
//     {ct0..ct3} {ct4, ct5, cb0, cb1} {cb2..cb5};  C_top(g) couples separator g - 1 with the first interior row of group g,
//     C_bot(g) the last interior row with separator g
// Delta^-1 is stored as its Cholesky factor W (Delta^-1 = W W^T, W lower triangular: w00 w10 w11 w20 w21 w22): the pose blocks are
// badly conditioned in world coordinates (the rotation couples to the translation with the lever arm |l| ~ the size of the
// world), so rounding the SIX entries of Delta^-1 to FP32 can cost positive definiteness, whereas W W^T is PSD whatever the rounding
template <typename S>
__device__ __forceinline__ void sym3_chol_to_float(const S a[6], float w[6]) {
    const double a00 = (double)a[0], a10 = (double)a[1], a20 = (double)a[2], a11 = (double)a[3], a21 = (double)a[4], a22 = (double)a[5];
    const double w00 = sqrt(fmax(a00, 1e-300));
    const double w10 = a10 / w00, w20 = a20 / w00;
    const double w11 = sqrt(fmax(a11 - w10 * w10, 1e-30 * fmax(a11, 1e-300)));
    const double w21 = (a21 - w20 * w10) / w11;
    const double w22 = sqrt(fmax(a22 - w20 * w20 - w21 * w21, 1e-30 * fmax(a22, 1e-300)));
    w[0] = (float)w00; w[1] = (float)w10; w[2] = (float)w11; w[3] = (float)w20; w[4] = (float)w21; w[5] = (float)w22;
}
__device__ __forceinline__ void chain_store_row(float* base, int sc, const float L[9], const float dv[6]) {
    base[0] = L[0]; base[1] = L[1]; base[2] = L[2]; base[3] = L[3];
    base[sc] = L[4]; base[sc + 1] = L[5]; base[sc + 2] = L[6]; base[sc + 3] = L[7];
    base[2 * sc] = L[8]; base[2 * sc + 1] = dv[0]; base[2 * sc + 2] = dv[1]; base[2 * sc + 3] = dv[2];
    base[3 * sc] = dv[3]; base[3 * sc + 1] = dv[4]; base[3 * sc + 2] = dv[5]; base[3 * sc + 3] = 0.f;
}
template <typename S>
__device__ __forceinline__ void sym6_to_full(const S a[6], S m[9]) {
    m[0] = a[0]; m[1] = a[1]; m[2] = a[2]; m[3] = a[1]; m[4] = a[3]; m[5] = a[4]; m[6] = a[2]; m[7] = a[4]; m[8] = a[5];
}
template <typename S>
__device__ __forceinline__ void mat3_mul(const S a[9], const S b[9], S o[9]) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
template <typename S>
__device__ __forceinline__ void mat3_tmul(const S a[9], const S b[9], S o[9]) {   // a^T b
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}

template <typename S>
__global__ void __launch_bounds__(128) k_pcg_chain_factor(Dev<S> d, PcgWork<S> w) {
    __shared__ S Eff[64][6], Ell[64][6], Efl[64][9], Dsep[64][6], Osep[64][9];
    // threads 0-63: group g's downward elimination (the stored factors) and the corner blocks that follow from it; threads 64-127: the
    // UPWARD elimination of the same group, which only yields [T^-1]_(0,0) and is independent of the first -- 61 instead of 91 sequential steps
    const bool upper = threadIdx.x >= 64;
    const int c = blockIdx.x, g = threadIdx.x & 63, cp = d.pc_cp, K = cp / 32, Kp = w.ch_Kp, cps = w.ch_cps;
    const size_t nrows = (size_t)d.pc_chunks * cp;
    float* F = w.chF + (size_t)c * w.ch_fac_floats;
    float* Fs = F + 16 * cps;
    float* Fc = Fs + 16 * Kp;
    const size_t R0 = (size_t)c * cp + (size_t)g * 32;
    auto ldD = [&](int k, S o[6]) {
#pragma unroll
        for (int q = 0; q < 6; q++) o[q] = w.chD[(size_t)q * nrows + R0 + k];
    };
    auto ldO = [&](long long k, S o[6]) {   // block between rows R0 + k and R0 + k + 1
#pragma unroll
        for (int q = 0; q < 6; q++) o[q] = w.chO[(size_t)q * nrows + R0 + k];
    };
    if (g < K && !upper) {
        // downward elimination of the interior: L_k = O_{k-1} Delta_{k-1}^-1, Delta_k = D_k - L_k O_{k-1}
        S dinv[6] = {S(0), S(0), S(0), S(0), S(0), S(0)}, op[9];
        S Dn[6], On[6];                      // row k's blocks are fetched one step ahead: the loads do not depend on the recurrence
        ldD(0, Dn); ldO(0, On);
        for (int k = 0; k < 31; k++) {
            S D[6], o6[6], L[9] = {S(0), S(0), S(0), S(0), S(0), S(0), S(0), S(0), S(0)};
#pragma unroll
            for (int q = 0; q < 6; q++) { D[q] = Dn[q]; o6[q] = On[q]; }
            if (k + 1 < 31) { ldD(k + 1, Dn); ldO(k + 1, On); }
            if (k > 0) {
                S di[9], lo[9];
                sym6_to_full<S>(dinv, di);
                mat3_mul<S>(op, di, L);
                mat3_mul<S>(L, op, lo);
                D[0] -= lo[0]; D[1] -= S(0.5) * (lo[1] + lo[3]); D[2] -= S(0.5) * (lo[2] + lo[6]);
                D[3] -= lo[4]; D[4] -= S(0.5) * (lo[5] + lo[7]); D[5] -= lo[8];
            }
            sym3_inverse<S>(D, dinv);
            {
                float Lf[9], df[6];
#pragma unroll
                for (int q = 0; q < 9; q++) Lf[q] = (float)L[q];
                sym3_chol_to_float<S>(dinv, df);
                chain_store_row(F + ((size_t)(k * 4) * Kp + g) * 4, Kp * 4, Lf, df);
            }
            sym6_to_full<S>(o6, op);
        }
        {   // the separator's slot of the interior arrays is never read by the solve
            const float z9[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            chain_store_row(F + ((size_t)(31 * 4) * Kp + g) * 4, Kp * 4, z9, z9);
        }
#pragma unroll
        for (int q = 0; q < 6; q++) Ell[g][q] = dinv[q];            // [T^-1]_(30,30)
        // [T^-1]_(0,30): X_30 = Delta_30^-1, X_k = -L_{k+1}^T X_{k+1} (with the stored factors)
        S X[9];
        sym6_to_full<S>(dinv, X);
        for (int k = 29; k >= 0; k--) {
            S L[9], t[9];
            const float* b = F + ((size_t)((k + 1) * 4) * Kp + g) * 4;
#pragma unroll
            for (int q = 0; q < 9; q++) L[q] = (S)b[(q >> 2) * Kp * 4 + (q & 3)];
            mat3_tmul<S>(L, X, t);
#pragma unroll
            for (int q = 0; q < 9; q++) X[q] = -t[q];
        }
#pragma unroll
        for (int q = 0; q < 9; q++) Efl[g][q] = X[q];
    }
    if (g < K && upper) {
        // upward elimination: Delta'_30 = D_30, Delta'_k = D_k - O_k Delta'_{k+1}^-1 O_k; [T^-1]_(0,0) = Delta'_0^-1
        S dinv[6];
        S up[6];
        ldD(30, up);
        sym3_inverse<S>(up, dinv);
        S Dn[6], On[6];
        ldD(29, Dn); ldO(29, On);
        for (int k = 29; k >= 0; k--) {
            S D[6], o6[6], o[9], di[9], t[9], lo[9];
#pragma unroll
            for (int q = 0; q < 6; q++) { D[q] = Dn[q]; o6[q] = On[q]; }
            if (k > 0) { ldD(k - 1, Dn); ldO(k - 1, On); }
            sym6_to_full<S>(o6, o); sym6_to_full<S>(dinv, di);
            mat3_mul<S>(o, di, t);
            mat3_mul<S>(t, o, lo);
            D[0] -= lo[0]; D[1] -= S(0.5) * (lo[1] + lo[3]); D[2] -= S(0.5) * (lo[2] + lo[6]);
            D[3] -= lo[4]; D[4] -= S(0.5) * (lo[5] + lo[7]); D[5] -= lo[8];
            sym3_inverse<S>(D, dinv);
        }
#pragma unroll
        for (int q = 0; q < 6; q++) Eff[g][q] = dinv[q];
    }
    __syncthreads();
    if (g < K && !upper) {
        // separator Schur system: eliminate the interiors on both sides of separator g
        S ct6[6] = {S(0), S(0), S(0), S(0), S(0), S(0)}, cb6[6], cn6[6] = {S(0), S(0), S(0), S(0), S(0), S(0)}, D[6];
        if (g > 0) ldO(-1, ct6);          // separator g-1 (row R0 - 1) <-> first interior row
        ldO(30, cb6);                      // last interior row <-> separator g
        if (g + 1 < K) ldO(31, cn6);       // separator g <-> first interior row of group g + 1  (= C_top(g + 1))
        ldD(31, D);
        S cb[9], ct[9], cn[9], e[9], t[9], u[9];
        sym6_to_full<S>(cb6, cb); sym6_to_full<S>(ct6, ct); sym6_to_full<S>(cn6, cn);
        sym6_to_full<S>(Ell[g], e);
        mat3_mul<S>(cb, e, t); mat3_mul<S>(t, cb, u);
        D[0] -= u[0]; D[1] -= S(0.5) * (u[1] + u[3]); D[2] -= S(0.5) * (u[2] + u[6]); D[3] -= u[4]; D[4] -= S(0.5) * (u[5] + u[7]); D[5] -= u[8];
        if (g + 1 < K) {
            sym6_to_full<S>(Eff[g + 1], e);
            mat3_mul<S>(cn, e, t); mat3_mul<S>(t, cn, u);
            D[0] -= u[0]; D[1] -= S(0.5) * (u[1] + u[3]); D[2] -= S(0.5) * (u[2] + u[6]); D[3] -= u[4]; D[4] -= S(0.5) * (u[5] + u[7]); D[5] -= u[8];
        }
#pragma unroll
        for (int q = 0; q < 6; q++) Dsep[g][q] = D[q];
        // block (separator g-1, separator g) = -C_top(g) [T_g^-1]_(0,30) C_bot(g)
        mat3_mul<S>(ct, Efl[g], t); mat3_mul<S>(t, cb, u);
#pragma unroll
        for (int q = 0; q < 9; q++) Osep[g][q] = -u[q];
#pragma unroll
        for (int q = 0; q < 12; q++) Fc[((q >> 2) * Kp + g) * 4 + (q & 3)] = (float)(q < 6 ? ct6[q] : cb6[q - 6]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // block LDL^T of the separator chain: L_j = Osep_j^T Delta_{j-1}^-1, Delta_j = Dsep_j - L_j Osep_j
        S dinv[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
        for (int j = 0; j < K; j++) {
            S D[6], L[9] = {S(0), S(0), S(0), S(0), S(0), S(0), S(0), S(0), S(0)};
#pragma unroll
            for (int q = 0; q < 6; q++) D[q] = Dsep[j][q];
            if (j > 0) {
                S di[9], lo[9];
                sym6_to_full<S>(dinv, di);
                mat3_tmul<S>(Osep[j], di, L);
                mat3_mul<S>(L, Osep[j], lo);
                D[0] -= lo[0]; D[1] -= S(0.5) * (lo[1] + lo[3]); D[2] -= S(0.5) * (lo[2] + lo[6]);
                D[3] -= lo[4]; D[4] -= S(0.5) * (lo[5] + lo[7]); D[5] -= lo[8];
            }
            sym3_inverse<S>(D, dinv);
            {
                float Lf[9], df[6];
#pragma unroll
                for (int q = 0; q < 9; q++) Lf[q] = (float)L[q];
                sym3_chol_to_float<S>(dinv, df);
                chain_store_row(Fs + 16 * j, 4, Lf, df);
            }
        }
    }
}

// One block-tridiagonal solve with stored factors: y_k = rhs_k - L_k y_{k-1}, w_k = Delta_k^-1 y_k, z_k = w_k - L_{k+1}^T z_{k+1}.
// rhs may alias out.  a0 / aN are added to the first / last right-hand side.  The recurrences run in FP32 like the stored factors
// (a preconditioner only has to approximate M^-1; the CG recurrences themselves stay in S).  One thread walks a chain, so the
// cost is the instruction count of a step: operands are float vectors in shared memory, the next step's are fetched while the
// current one computes (two register sets, no copies).
//   F: the first row's float4 group; fc4 = distance between a row's four float4 (in float4), fk4 = distance between rows
#define BOS_CH_FETCH(P, FP, RP)                                                                       \
    P##a = (FP)[0]; P##b = (FP)[fc4]; P##c = (FP)[2 * fc4]; P##d = (FP)[3 * fc4];                        \
    P##r0 = (RP)[0]; P##r1 = (RP)[vcs]; P##r2 = (RP)[2 * vcs];
#define BOS_CH_FWD(P)                                                                                  \
    {                                                                                                  \
        const float n0 = fmaf(-P##a.z, y2, fmaf(-P##a.y, y1, fmaf(-P##a.x, y0, P##r0)));                  \
        const float n1 = fmaf(-P##b.y, y2, fmaf(-P##b.x, y1, fmaf(-P##a.w, y0, P##r1)));                  \
        const float n2 = fmaf(-P##c.x, y2, fmaf(-P##b.w, y1, fmaf(-P##b.z, y0, P##r2)));                  \
        y0 = n0; y1 = n1; y2 = n2;                                                                     \
        const float t0 = P##c.y * y0 + P##c.z * y1 + P##d.x * y2;         /* W^T y, W = (c.y; c.z c.w; d.x d.y d.z) */ \
        const float t1 = P##c.w * y1 + P##d.y * y2;                                                        \
        const float t2 = P##d.z * y2;                                                                      \
        op[0] = P##c.y * t0;                                                                               \
        op[vcs] = P##c.z * t0 + P##c.w * t1;                                                                \
        op[2 * vcs] = P##d.x * t0 + P##d.y * t1 + P##d.z * t2;                                              \
        op += vstep;                                                                                   \
    }
#define BOS_CH_BFETCH(P, FP, OP)                                                                      \
    P##a = (FP)[0]; P##b = (FP)[fc4]; P##e = (FP)[2 * fc4].x; P##w0 = (OP)[0]; P##w1 = (OP)[vcs]; P##w2 = (OP)[2 * vcs];
#define BOS_CH_BWD(P)                                                                                  \
    {                                                                                                  \
        const float n0 = fmaf(-P##b.z, z2, fmaf(-P##a.w, z1, fmaf(-P##a.x, z0, P##w0)));                  \
        const float n1 = fmaf(-P##b.w, z2, fmaf(-P##b.x, z1, fmaf(-P##a.y, z0, P##w1)));                  \
        const float n2 = fmaf(-P##e, z2, fmaf(-P##b.y, z1, fmaf(-P##a.z, z0, P##w2)));                    \
        z0 = n0; z1 = n1; z2 = n2;                                                                     \
        op[0] = z0; op[vcs] = z1; op[2 * vcs] = z2;                                                    \
    }
__device__ __forceinline__ void chain_thomas(const float4* F, int fc4, int fk4, const float* rhs, float* out, int vcs, int vstep, int n, const float a0[3],
                                             const float aN[3]) {
    float y0 = 0.f, y1 = 0.f, y2 = 0.f;
    float4 Aa, Ab, Ac, Ad, Ba, Bb, Bc, Bd;
    float Ar0, Ar1, Ar2, Br0, Br1, Br2;
    const float4* Fp = F;
    const float* rp = rhs;
    float* op = out;
    BOS_CH_FETCH(A, Fp, rp)
    Ar0 += a0[0]; Ar1 += a0[1]; Ar2 += a0[2];
    if (n == 1) { Ar0 += aN[0]; Ar1 += aN[1]; Ar2 += aN[2]; }
    for (int k = 0;;) {   // the fetch past the last row re-reads the last row (unused)
        if (k + 1 < n) { Fp += fk4; rp += vstep; }
        BOS_CH_FETCH(B, Fp, rp)
        if (k + 2 == n) { Br0 += aN[0]; Br1 += aN[1]; Br2 += aN[2]; }
        BOS_CH_FWD(A)
        if (++k >= n) break;
        if (k + 1 < n) { Fp += fk4; rp += vstep; }
        BOS_CH_FETCH(A, Fp, rp)
        if (k + 2 == n) { Ar0 += aN[0]; Ar1 += aN[1]; Ar2 += aN[2]; }
        BOS_CH_FWD(B)
        if (++k >= n) break;
    }
    // backward: op is one past the last row, Fp at the last row; row k needs L_{k+1} (zero above the last row) and w_k
    op -= vstep;
    float z0 = 0.f, z1 = 0.f, z2 = 0.f;
    float4 Pa = make_float4(0.f, 0.f, 0.f, 0.f), Pb = Pa, Qa, Qb;
    float Pe = 0.f, Pw0 = op[0], Pw1 = op[vcs], Pw2 = op[2 * vcs], Qe, Qw0, Qw1, Qw2;
    for (int k = n - 1;;) {   // row k - 1 needs L_k (at Fp) and its own w; at k = 0 the fetch is unused
        { const float* wp = (k > 0) ? op - vstep : op; BOS_CH_BFETCH(Q, Fp, wp) }
        BOS_CH_BWD(P)
        if (--k < 0) break;
        op -= vstep; Fp -= fk4;
        { const float* wp = (k > 0) ? op - vstep : op; BOS_CH_BFETCH(P, Fp, wp) }
        BOS_CH_BWD(Q)
        if (--k < 0) break;
        op -= vstep; Fp -= fk4;
    }
}

// z = M^-1 r for the CTA's chunk; called by threads 0..63 (two warps, named barrier 1).  All vectors are FP32 [3][cps] in the
// transposed row layout (row (g, k) at k * Kp + g) at byte offsets of the dynamic shared memory: r1 / w1 right-hand side and work
// space of the first interior pass and of the separator system (may coincide), r2 / w2 of the second pass (w2 holds the result;
// r2 == w2 allowed).  Not inlined: inside the persistent kernel the 64-register budget is taken by the per-thread row state, here
// the recurrences get their own allocation.
__device__ __noinline__ void chain_apply(unsigned fac_off, unsigned r1_off, unsigned w1_off, unsigned r2_off, unsigned w2_off, int cps, int Kp, int K) {
    extern __shared__ __align__(16) unsigned char pcg_smem[];
    const float* r1 = reinterpret_cast<const float*>(pcg_smem + r1_off);
    float* w1 = reinterpret_cast<float*>(pcg_smem + w1_off);
    const float* r2 = reinterpret_cast<const float*>(pcg_smem + r2_off);
    float* w2 = reinterpret_cast<float*>(pcg_smem + w2_off);
    const int g = threadIdx.x;
    const bool act = g < K;
    const float4* F4 = reinterpret_cast<const float4*>(pcg_smem + fac_off);
    const float4* Fs4 = F4 + 4 * cps;          // 16 * cps floats
    const float4* Fc4 = Fs4 + 4 * Kp;          // 16 * Kp floats further
    const float zero3[3] = {0.f, 0.f, 0.f};
    float ct[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, cb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int sp = 31 * Kp + g;
    if (act) {
        const float4 c0 = Fc4[g], c1 = Fc4[Kp + g], c2 = Fc4[2 * Kp + g];
        ct[0] = c0.x; ct[1] = c0.y; ct[2] = c0.z; ct[3] = c0.w; ct[4] = c1.x; ct[5] = c1.y;
        cb[0] = c1.z; cb[1] = c1.w; cb[2] = c2.x; cb[3] = c2.y; cb[4] = c2.z; cb[5] = c2.w;
        chain_thomas(F4 + g, Kp, 4 * Kp, r1 + g, w1 + g, cps, Kp, 31, zero3, zero3);
    }
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (act) {   // separator right-hand side: r_sep - C_bot(g) y_(g,30) - C_top(g+1) y_(g+1,0)
        const int lp = 30 * Kp + g;
        float q0 = r1[sp], q1 = r1[sp + cps], q2 = r1[sp + 2 * cps];
        {
            const float v0 = w1[lp], v1 = w1[lp + cps], v2 = w1[lp + 2 * cps];
            q0 -= cb[0] * v0 + cb[1] * v1 + cb[2] * v2; q1 -= cb[1] * v0 + cb[3] * v1 + cb[4] * v2; q2 -= cb[2] * v0 + cb[4] * v1 + cb[5] * v2;
        }
        if (g + 1 < K) {
            const int q = g + 1;
            const float4 c0 = Fc4[q], c1 = Fc4[Kp + q];
            const float v0 = w1[q], v1 = w1[q + cps], v2 = w1[q + 2 * cps];
            q0 -= c0.x * v0 + c0.y * v1 + c0.z * v2; q1 -= c0.y * v0 + c0.w * v1 + c1.x * v2; q2 -= c0.z * v0 + c1.x * v1 + c1.y * v2;
        }
        w1[sp] = q0; w1[sp + cps] = q1; w1[sp + 2 * cps] = q2;
    }
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (g == 0) chain_thomas(Fs4, 1, 4, w1 + 31 * Kp, w1 + 31 * Kp, cps, 1, K, zero3, zero3);
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (act) {   // interiors again, with the separator solutions moved to the right-hand side
        float a0[3] = {0.f, 0.f, 0.f}, aN[3];
        if (g > 0) {
            const float v0 = w1[sp - 1], v1 = w1[sp - 1 + cps], v2 = w1[sp - 1 + 2 * cps];
            a0[0] = -(ct[0] * v0 + ct[1] * v1 + ct[2] * v2); a0[1] = -(ct[1] * v0 + ct[3] * v1 + ct[4] * v2); a0[2] = -(ct[2] * v0 + ct[4] * v1 + ct[5] * v2);
        }
        const float v0 = w1[sp], v1 = w1[sp + cps], v2 = w1[sp + 2 * cps];
        aN[0] = -(cb[0] * v0 + cb[1] * v1 + cb[2] * v2); aN[1] = -(cb[1] * v0 + cb[3] * v1 + cb[4] * v2); aN[2] = -(cb[2] * v0 + cb[4] * v1 + cb[5] * v2);
        chain_thomas(F4 + g, Kp, 4 * Kp, r2 + g, w2 + g, cps, Kp, 31, a0, aN);
        w2[sp] = v0; w2[sp + cps] = v1; w2[sp + 2 * cps] = v2;   // the separator's own solution joins the result
    }
}

// ---- coarse space of the chain preconditioner ---------------------------------------------------------------------------
// The chunk-exact chain solve leaves the smooth, chunk-spanning error modes to CG.  They are taken by a coarse space of
// piecewise-linear hats along the pose order: node c sits at the start of chunk c (pc_chunks + 1 nodes, 3 dof each); row r of
// chunk c carries the weights 1 - t and t, t = (r + 1/2) / cp, for nodes c and c + 1 (zero for the fixed pose).  The coarse
// operator is the Galerkin product A_c = P^T S P with the TRUE Schur complement, so it also sees the pose-pose coupling through
// the landmarks; z = M_chunk^-1 r + P A_c^-1 P^T r (two-level additive Schwarz, SPD).  A_c is assembled once per solve:
//   P^T Hpp P                      k_coarse_pose  (per chunk in registers, one set of atomics per chunk)
//   - sum_l G_l^T Hll_l^-1 G_l     k_coarse_lm    (one thread per landmark walks its edges)
// then factorised (dense_cholesky_lower) and inverted explicitly (k_coarse_inverse): applying it is a 6-row mat-vec per CTA.
// Node geometry: every chunk is cut into c_nseg segments of c_h = 32 m rows (the last one may be shorter); node n sits at the start of
// GLOBAL segment n = c * c_nseg + j, one more node closes the last segment.  Row r of a chunk carries the weights 1 - t and t for the nodes
// at the two ends of its segment, t = (r - j h + 1/2) / len_j.  (One segment per chunk = the round-1 layout; four per chunk cut the CG
// iterations at synth-2M from 73 to 46: the nodes then also resolve the variation ALONG a serpentine row, tests/precond_model.py.)
__host__ __device__ __forceinline__ void coarse_seg(int r, int cp, int h, int& j, float& t) {
    j = r / h;
    const int len = (cp - j * h < h) ? cp - j * h : h;
    t = ((float)(r - j * h) + 0.5f) / (float)len;
}

// one thread per compact landmark row.  Its edges come in ascending pose order (sliced-ELL row), so the chunk index never
// decreases: the edges of one chunk are summed into G (2 x 3 per node: rows = the two components of Jl, weighted by the hat
// weights of the pose), neighbouring chunks share a node; then the lower triangle of -G^T Hll^-1 G goes into A_c.  A landmark
// seen from more than kCoarseMaxChunks chunks is left out altogether (A_c only grows: still SPD).
constexpr int kCoarseMaxChunks = 8;     // ... segments, since the nodes sit at segment ends
constexpr int kCoarseMaxSeg = 8;        // segments per chunk, at most
template <typename S>
__global__ void __launch_bounds__(128) k_coarse_lm(Dev<S> d, PcgWork<S> w) {
    constexpr int RPG = 32 / kEllLanesL;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= d.n_clm) return;
    const int g = row / RPG, lane0 = (row % RPG) * kEllLanesL, cp = d.pc_cp, hseg = w.c_h, nseg = w.c_nseg;
    const int off = __ldg(d.ell_Loff + g), W = __ldg(d.ell_Loff + g + 1) - off;
    const S lx = w.ul4[4LL * row + 2], ly = w.ul4[4LL * row + 3];
    const S so_u = (S)w.sqrt_omega;
    int nid[2 * kCoarseMaxChunks] = {0};
    double G[2 * kCoarseMaxChunks][6];
    int nn = 0, cur = -1;
    bool over = false;
    double acc[12];
    auto flush = [&]() {
        if (cur < 0) return;
        if (nn + 2 > 2 * kCoarseMaxChunks) { over = true; return; }
        int a;
        if (nn > 0 && nid[nn - 1] == cur) a = nn - 1;
        else { a = nn++; nid[a] = cur; for (int e = 0; e < 6; e++) G[a][e] = 0.0; }
        for (int e = 0; e < 6; e++) G[a][e] += acc[e];
        const int b = nn++;
        nid[b] = cur + 1;
        for (int e = 0; e < 6; e++) G[b][e] = acc[6 + e];
    };
    for (int tb = 0; tb < W; tb++)
        for (int q = 0; q < kEllLanesL; q++) {
            const long long slot = ((long long)off + tb) * 32 + lane0 + q;
            const int ps = __ldg(d.ell_Lpose + slot);
            if (ps < 0) continue;                                   // padding, or an edge of the fixed pose
            const int c = ps / cp;
            int jseg; float tf;
            coarse_seg(ps - c * cp, cp, hseg, jseg, tf);
            const int gs = c * nseg + jseg;
            if (gs != cur) {
                flush();
                cur = gs;
                for (int e = 0; e < 12; e++) acc[e] = 0.0;
            }
            S px, py;
            load_lm<S>(d.pose, 2 * ps, px, py);
            S j0, j1;
            bearing_jl_world<S>(px, py, lx, ly, j0, j1);
            const S so = w.omega_uniform ? so_u : __ldg(w.Lw + slot);
            j0 *= so; j1 *= so;
            const double jp[3] = {(double)-j0, (double)-j1, (double)(j0 * ly - j1 * lx)};
            const double t = (double)tf, wl = 1.0 - t;
            for (int dd = 0; dd < 3; dd++) {
                const double a0 = (double)j0 * jp[dd], a1 = (double)j1 * jp[dd];
                acc[dd] += wl * a0; acc[3 + dd] += wl * a1; acc[6 + dd] += t * a0; acc[9 + dd] += t * a1;
            }
        }
    flush();
    if (over || nn == 0) return;
    const double i00 = (double)w.hllinv_c[3LL * row], i01 = (double)w.hllinv_c[3LL * row + 1], i11 = (double)w.hllinv_c[3LL * row + 2];
    const int nc = w.c_nc;
    for (int a = 0; a < nn; a++) {
        double T0[3], T1[3];   // Hll^-1 G[a]
        for (int e = 0; e < 3; e++) { T0[e] = i00 * G[a][e] + i01 * G[a][3 + e]; T1[e] = i01 * G[a][e] + i11 * G[a][3 + e]; }
        for (int b = 0; b <= a; b++)
            for (int dd = 0; dd < 3; dd++)
                for (int e = 0; e < 3; e++) {
                    if (a == b && e > dd) continue;
                    // row (node a, dd), column (node b, e): G[a](:, dd)^T Hll^-1 G[b](:, e)
                    const double v = T0[dd] * G[b][e] + T1[dd] * G[b][3 + e];
                    atomicAdd(w.cA + (size_t)(3 * nid[a] + dd) + (size_t)(3 * nid[b] + e) * nc, -v);
                }
    }
}

// P^T Hpp P: one CTA per chunk.  Terms whose two poses lie in this chunk are summed in registers (lower triangle of the 6x6
// block of nodes c, c + 1), the few that reach into another chunk (chunk-boundary odometry edges, loop closures) go straight
// to A_c.  Every ordered pose pair (i, j) is visited once from i; a term is kept iff it lands on or below the diagonal.
template <typename S>
__global__ void __launch_bounds__(256) k_coarse_pose(Dev<S> d, PcgWork<S> w) {
    __shared__ double red[8][21];
    const int cp = d.pc_cp, nc = w.c_nc, hseg = w.c_h, nseg = w.c_nseg;
    const int c = blockIdx.x / nseg, jseg = blockIdx.x % nseg, gs = blockIdx.x;     // one CTA per (chunk, segment)
    const int r_lo = jseg * hseg, r_hi = (r_lo + hseg < cp) ? r_lo + hseg : cp;
    double acc[21];
#pragma unroll
    for (int q = 0; q < 21; q++) acc[q] = 0.0;
    for (int r = r_lo + (int)threadIdx.x; r < r_hi; r += blockDim.x) {
        const int i = __ldg(d.pc_row_pose + (size_t)c * cp + r);
        if (i < 0 || i == d.fixed) continue;
        int jj; float t;
        coarse_seg(r, cp, hseg, jj, t);
        const double wi[2] = {1.0 - (double)t, (double)t};
        double H[9];
        {
            const S* hp = d.Hpp + 6LL * i;
            H[0] = hp[0]; H[1] = hp[1]; H[2] = hp[2]; H[3] = hp[1]; H[4] = hp[3]; H[5] = hp[4]; H[6] = hp[2]; H[7] = hp[4]; H[8] = hp[5];
        }
#pragma unroll
        for (int sa = 0; sa < 2; sa++)
#pragma unroll
            for (int sb = 0; sb < 2; sb++)
#pragma unroll
                for (int dd = 0; dd < 3; dd++)
#pragma unroll
                    for (int e = 0; e < 3; e++) {
                        const int row = 3 * sa + dd, col = 3 * sb + e;
                        if (row >= col) acc[row * (row + 1) / 2 + col] += wi[sa] * wi[sb] * H[3 * dd + e];
                    }
        for (int q = __ldg(d.pp_ptr + i); q < __ldg(d.pp_ptr + i + 1); q++) {
            const int j = __ldg(d.pp_nbr + q);
            if (j == d.fixed) continue;
            const int sl = __ldg(d.pp_slot + q);
            const S* Bo = d.Hoff + 9LL * (sl & 0x7fffffff);
            // H_ij: the stored block is H[lo][hi]; sl >= 0 means i is the lo side
#pragma unroll
            for (int dd = 0; dd < 3; dd++)
#pragma unroll
                for (int e = 0; e < 3; e++) H[3 * dd + e] = (sl >= 0) ? (double)Bo[3 * dd + e] : (double)Bo[3 * e + dd];
            const int cj = j / cp;
            int js; float tj;
            coarse_seg(j - cj * cp, cp, hseg, js, tj);
            const int gsj = cj * nseg + js;
            const double wj[2] = {1.0 - (double)tj, (double)tj};
            if (gsj == gs) {
#pragma unroll
                for (int sa = 0; sa < 2; sa++)
#pragma unroll
                    for (int sb = 0; sb < 2; sb++)
#pragma unroll
                        for (int dd = 0; dd < 3; dd++)
#pragma unroll
                            for (int e = 0; e < 3; e++) {
                                const int row = 3 * sa + dd, col = 3 * sb + e;
                                if (row >= col) acc[row * (row + 1) / 2 + col] += wi[sa] * wj[sb] * H[3 * dd + e];
                            }
            } else {
                for (int sa = 0; sa < 2; sa++)
                    for (int sb = 0; sb < 2; sb++)
                        for (int dd = 0; dd < 3; dd++)
                            for (int e = 0; e < 3; e++) {
                                const int row = 3 * (gs + sa) + dd, col = 3 * (gsj + sb) + e;
                                if (row >= col) atomicAdd(w.cA + (size_t)row + (size_t)col * nc, wi[sa] * wj[sb] * H[3 * dd + e]);
                            }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 21; q++) {
        const double v = warp_sum(acc[q]);
        if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 21) {
        double v = 0.0;
        for (int k = 0; k < 8; k++) v += red[k][threadIdx.x];
        int row = 0;
        while ((row + 1) * (row + 2) / 2 <= (int)threadIdx.x) row++;
        const int col = threadIdx.x - row * (row + 1) / 2;
        if (v != 0.0) atomicAdd(w.cA + (size_t)(3 * gs + row) + (size_t)(3 * gs + col) * nc, v);
    }
}

// unsupported coarse dofs (empty chunks) get a unit diagonal: their residual is always zero
__global__ void k_coarse_fix(double* A, int nc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nc && !(A[(size_t)j * nc + j] > 0.0)) A[(size_t)j * nc + j] = 1.0;
}

// ---- banded factorisation / inverse of the coarse operator ---------------------------------------------------------------------
// A_c couples node n only with nodes a few segments away (the observers of a landmark are close to each other along the pose order, pose-pose
// blocks join consecutive poses): it is a band matrix of half bandwidth bw, and its Cholesky factor keeps the band.  ONE CTA factorises it with
// the active (bw + 1) x (bw + 1) window in shared memory (a ring over the columns); finished columns go to two compact copies of the band with row
// stride W = bw + 1:  Lc[k][e] = L[k + e][k] (column k, contiguous) and Lr[i][e] = L[i][i - e] (row i, contiguous), e = 0..bw, plus Ldi[k] = 1 / L[k][k].
// (Round 1 ran 21 dense-Cholesky launches + a dense triangular inverse here: 1.2 ms per solve for a 447 x 447 matrix.)
constexpr int kBandCholThreads = 256;      // a 16 x 16 grid of threads over the (a, b) pairs of the trailing window
__global__ void __launch_bounds__(kBandCholThreads) k_coarse_band_chol(const double* __restrict__ A, int nc, int bw, double* __restrict__ Lc,
                                                                       double* __restrict__ Lr, double* __restrict__ Ldi, double* __restrict__ stats) {
    extern __shared__ double win[];           // [R][W]: slot (k % R), entry e = row k + e.  R = bw + 2: one slot more than the window, so that the
                                              // column entering the ring is not yet part of the NEXT step's trailing window (one barrier per step)
    const int W = bw + 1, R = bw + 2, tid = threadIdx.x, ta = tid >> 4, tb = tid & 15;
    for (int q = tid; q < R * W; q += kBandCholThreads) {
        const int k = q / W, e = q - k * W;
        win[q] = (k < nc && k + e < nc) ? A[(size_t)(k + e) + (size_t)k * nc] : 0.0;
    }
    // column j + R enters the ring when column j leaves it; its entries travel through registers, loaded one column step early so that the
    // global-memory latency is hidden behind a whole step (thread e holds entry e; W <= 160 < 256 threads)
    double pre = 0.0;
    auto prefetch = [&](int kn) { pre = (tid < W && kn < nc && kn + tid < nc) ? __ldg(A + (size_t)(kn + tid) + (size_t)kn * nc) : 0.0; };
    prefetch(R);
    __syncthreads();
    bool bad = false;
    int sj = 0;                                // j % R
    for (int j = 0; j < nc; j++, sj = (sj + 1 == R) ? 0 : sj + 1) {
        double* cj = win + (size_t)sj * W;
        // every thread derives the pivot itself: no barrier between the pivot and the update, which works on the UNSCALED column
        double p = cj[0];
        if (!(p > 0.0)) { bad = true; p = 1.0; }
        const double ip = 1.0 / p, rs = sqrt(ip);
        const double held = pre;               // column j + R, loaded during step j - 1
        prefetch(j + 1 + R);
        // trailing window: column j + a (a = 1..bw), row j + b (b = a..bw):  A[j+b][j+a] -= A[j+b][j] A[j+a][j] / A[j][j]
        for (int a = 1 + ta; a <= bw; a += 16) {
            const double la = cj[a] * ip;
            const int sa = (sj + a >= R) ? sj + a - R : sj + a;
            double* ca = win + (size_t)sa * W - a;
            for (int b = 1 + tb + ((a - 1 - tb + 15) >> 4 << 4); b <= bw; b += 16)      // first b >= a with b = 1 + tb (mod 16)
                if (j + b < nc) ca[b] -= cj[b] * la;
        }
        // the finished column: L[j + e][j] = A[j + e][j] / sqrt(A[j][j])
        if (tid < W) {
            const int e = tid;
            const double v = (j + e < nc) ? ((e == 0) ? p * rs : cj[e] * rs) : 0.0;
            Lc[(size_t)j * W + e] = v;
            if (j + e < nc) Lr[(size_t)(j + e) * W + e] = v;
            if (e == 0) Ldi[j] = rs;
        }
        __syncthreads();
        if (tid < W) cj[tid] = held;           // the slot of column j now holds column j + R: first touched two steps later, after the next barrier
    }
    if (bad && tid == 0) stats[5] = 1.0;       // the persistent kernel drops the coarse term for this solve
}
// rows of Lr above the band's start were never written: they are read only where i - e >= 0
// A_c^-1 column by column from the banded factor: one warp per column j.  L y = e_j by column sweeps starting at k = j, then L^T x = y by column
// sweeps of L^T (= rows of L) from the bottom; the vector lives in shared memory, the next band column is in registers before the current step ends.
constexpr int kBandMaxPerLane = 5;       // bw <= 159
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_coarse_band_inverse(const double* __restrict__ Lc, const double* __restrict__ Lr, const double* __restrict__ Ldi,
                                                                    double* __restrict__ Ainv, int nc, int bw, int ld) {
    extern __shared__ double vsh[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = bw + 1;
    const int j = blockIdx.x * WARPS + warp;
    if (j >= nc) return;
    double* v = vsh + (size_t)warp * nc;
    for (int i = lane; i < nc; i += 32) v[i] = (i == j) ? 1.0 : 0.0;
    __syncwarp();
    double cur[kBandMaxPerLane], nxt[kBandMaxPerLane], dcur, dnxt;
    auto fetch = [&](const double* base, int k, double* dst, double& dg) {      // entries e = 1 + lane + 32 q of band line k, and 1 / L[k][k]
        const bool in = k >= 0 && k < nc;
#pragma unroll
        for (int q = 0; q < kBandMaxPerLane; q++) {
            const int e = 1 + lane + 32 * q;
            dst[q] = (in && e <= bw) ? __ldg(base + (size_t)k * W + e) : 0.0;
        }
        dg = in ? __ldg(Ldi + k) : 0.0;
    };
    fetch(Lc, j, cur, dcur);
    for (int k = j; k < nc; k++) {
        fetch(Lc, k + 1, nxt, dnxt);
        const double yk = v[k] * dcur;
        __syncwarp();
        if (lane == 0) v[k] = yk;
#pragma unroll
        for (int q = 0; q < kBandMaxPerLane; q++) {
            const int e = 1 + lane + 32 * q;
            if (e <= bw && k + e < nc) v[k + e] -= cur[q] * yk;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kBandMaxPerLane; q++) cur[q] = nxt[q];
        dcur = dnxt;
    }
    fetch(Lr, nc - 1, cur, dcur);
    for (int i = nc - 1; i >= 0; i--) {
        fetch(Lr, i - 1, nxt, dnxt);
        const double xi = v[i] * dcur;
        __syncwarp();
        if (lane == 0) v[i] = xi;
#pragma unroll
        for (int q = 0; q < kBandMaxPerLane; q++) {
            const int e = 1 + lane + 32 * q;
            if (e <= bw && i - e >= 0) v[i - e] -= cur[q] * xi;      // Lr[i][e] = L[i][i - e] = (L^T)[i - e][i]
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kBandMaxPerLane; q++) cur[q] = nxt[q];
        dcur = dnxt;
    }
    for (int i = lane; i < nc; i += 32) Ainv[(size_t)j * ld + i] = v[i];
}

// explicit inverse from the Cholesky factor (column-major lower): one warp per column j, L y = e_j by column sweeps, L^T x = y
// by dot products; the y / x vector lives in shared memory, the next column of L is in registers before the current step ends
constexpr int kCoarseMaxPerLane = 15;   // ceil(3 * 160 / 32)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_coarse_inverse(const double* __restrict__ L, double* __restrict__ Ainv, int nc, int ld) {
    extern __shared__ double vsh[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * WARPS + warp;
    if (j >= nc) return;
    double* v = vsh + (size_t)warp * nc;
    for (int i = lane; i < nc; i += 32) v[i] = (i == j) ? 1.0 : 0.0;
    __syncwarp();
    double col[kCoarseMaxPerLane], nxt[kCoarseMaxPerLane];
    auto fetch = [&](int k, double* dst) {   // column k from its diagonal down: entry k + lane + 32 q
#pragma unroll
        for (int q = 0; q < kCoarseMaxPerLane; q++) {
            const int i = k + lane + 32 * q;
            dst[q] = (k < nc && i < nc) ? __ldg(L + (size_t)i + (size_t)k * nc) : 0.0;   // predicated off beyond the column's end
        }
    };
    fetch(j, col);
    for (int k = j; k < nc; k++) {
        fetch(k + 1, nxt);
        const double dkk = __shfl_sync(0xffffffffu, col[0], 0);
        const double yk = v[k] / dkk;
        const int cnt = (nc - k + 31) / 32;                       // live entries per lane (warp-uniform)
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kCoarseMaxPerLane; q++) {
            if (q < cnt) {
                const int i = k + lane + 32 * q;
                if (i < nc) v[i] = (i == k) ? yk : v[i] - col[q] * yk;
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kCoarseMaxPerLane; q++) col[q] = nxt[q];
    }
    fetch(nc - 1, col);
    for (int i = nc - 1; i >= 0; i--) {
        fetch(i - 1 >= 0 ? i - 1 : nc, nxt);
        const int cnt = (nc - i + 31) / 32;
        double sacc = 0.0;
#pragma unroll
        for (int q = 0; q < kCoarseMaxPerLane; q++) {
            if (q < cnt) {
                const int k = i + lane + 32 * q;
                if (k > i && k < nc) sacc += col[q] * v[k];
            }
        }
        sacc = warp_sum(sacc);
        const double dii = __shfl_sync(0xffffffffu, col[0], 0);
        if (lane == 0) v[i] = (v[i] - sacc) / dii;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kCoarseMaxPerLane; q++) col[q] = nxt[q];
    }
    for (int i = lane; i < nc; i += 32) Ainv[(size_t)j * ld + i] = v[i];
}

// Coarse solve of one chunk, run by warps 2 .. 2 + kCoarseWarps - 1 of the persistent kernel WHILE warps 0-1 do the chunk's chain solve (the two halves
// of the preconditioner are independent until they are added): wait for the exchange of P^T r, then x_c = A_c^-1 rc at this chunk's nr <= 15
// node scalars.  A_c^-1 is symmetric, so lane m reads the nr CONTIGUOUS entries A_c^-1[m][row0 .. row0 + nr) of "its" columns m and the two halves
// of rc[m] straight from global memory: no shared-memory copy of rc (the factors occupy that space while the chain solve runs).
constexpr int kCoarseWarps = 24;     // 24 warps x 16 row slots = the 384 doubles of the reduction scratch
constexpr int kCoarseNr = 15;        // 3 * (4 segments + 1) <= 16 row slots
__device__ __noinline__ void coarse_apply(const double* __restrict__ Ainv, int ld, const double* __restrict__ cRc, int nc, int nsegs_all, int row0, int nr,
                                          const unsigned* counter, unsigned target, double* part /* [kCoarseWarps][16] */, double* xc) {
    const int t = threadIdx.x - 64, lane = t & 31, wq = t >> 5;
    if (t == 0)
        while (ld_acquire_u32(counter) < target) { }
    asm volatile("bar.sync 2, %0;" ::"n"(kCoarseWarps * 32) : "memory");
    // a HALF-WARP per column m: lane j of it reads A_c^-1[m][row0 + j] (the 16 lanes read 128 contiguous bytes) and the two halves of rc[m]
    // (one address for the whole half-warp); 48 half-warps stride over the columns, four columns in flight
    const int j = lane & 15, hw = 2 * wq + (lane >> 4);
    const bool live = j < nr;
    auto rc_of = [&](int m) {
        const int nd = m / 3, a = m - 3 * nd;
        double rv = 0.0;
        if (nd < nsegs_all) rv += __ldcg(cRc + 6LL * nd + a);
        if (nd > 0) rv += __ldcg(cRc + 6LL * (nd - 1) + 3 + a);
        return rv;
    };
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    constexpr int kStride = 2 * kCoarseWarps;
    int m = hw;
    for (; m + 3 * kStride < nc; m += 4 * kStride) {
        const double r0 = rc_of(m), r1 = rc_of(m + kStride), r2 = rc_of(m + 2 * kStride), r3 = rc_of(m + 3 * kStride);
        const double v0 = live ? __ldg(Ainv + (size_t)m * ld + row0 + j) : 0.0, v1 = live ? __ldg(Ainv + (size_t)(m + kStride) * ld + row0 + j) : 0.0;
        const double v2 = live ? __ldg(Ainv + (size_t)(m + 2 * kStride) * ld + row0 + j) : 0.0, v3 = live ? __ldg(Ainv + (size_t)(m + 3 * kStride) * ld + row0 + j) : 0.0;
        a0 += v0 * r0; a1 += v1 * r1; a2 += v2 * r2; a3 += v3 * r3;
    }
    for (; m < nc; m += kStride) a0 += (live ? __ldg(Ainv + (size_t)m * ld + row0 + j) : 0.0) * rc_of(m);
    double sv = (a0 + a1) + (a2 + a3);
    sv += __shfl_xor_sync(BOS_FULL_MASK, sv, 16);            // the warp's two half-warps
    if (lane < 16) part[wq * 16 + lane] = sv;
    asm volatile("bar.sync 2, %0;" ::"n"(kCoarseWarps * 32) : "memory");
    if (t < nr) {
        double tot = 0.0;
        for (int q = 0; q < kCoarseWarps; q++) tot += part[q * 16 + t];
        xc[t] = tot;
    }
}

// shared-memory plan of the persistent kernel (per CTA = per chunk)
template <typename S>
struct PcgSmemPlan {
    size_t vec_off, rec_off, loc_off, bytes;
    int cps, Kp, fac_floats;
    __host__ __device__ PcgSmemPlan(int cp, int cl_max, int slots_max, bool chain, int nc_coarse = 0) {
        Kp = (cp / 32) | 1;
        cps = chain ? Kp * 32 : cp;                                    // chain: rows transposed to (position in group, group), odd group stride
        fac_floats = 16 * Kp * 32 + 28 * Kp;
        vec_off = 0;                                                   // [12][cps]  p 0-2, s 3-5, r 6-8, yoff 9-11 (chain: then z = M^-1 r)
        rec_off = vec_off + (size_t)12 * cps * sizeof(S);              // [cl_max][4]  u0, u1, lx, ly of the chunk's landmarks
        size_t rec_bytes = (size_t)4 * (cl_max > 0 ? cl_max : 1) * sizeof(S);
        if (rec_bytes < (size_t)2 * cp * sizeof(S)) rec_bytes = (size_t)2 * cp * sizeof(S);   // ... and, before it is staged, the chunk's pose positions [cp][2]
        loc_off = rec_off + rec_bytes;                                 // [slots_max] 16-bit landmark table indices
        bytes = (loc_off + (size_t)2 * (slots_max > 0 ? slots_max : 1) + 15) / 16 * 16;
        // chain: the FP32 factors are staged over the record / index region while the preconditioner runs (the indices are re-staged after)
        if (chain && rec_off + (size_t)fac_floats * 4 > bytes) bytes = rec_off + (size_t)fac_floats * 4;
        // ... and, once the chunk solve is done, by the coarse residual (nc_coarse doubles)
        if (chain && rec_off + (size_t)nc_coarse * 8 > bytes) bytes = (rec_off + (size_t)nc_coarse * 8 + 15) / 16 * 16;
    }
};

template <typename S>
__global__ void __launch_bounds__(kPcgThreads, 1) k_pcg_fused(Dev<S> d, PcgWork<S> w, int max_iters, double tol2) {
    extern __shared__ __align__(16) unsigned char pcg_smem[];
    __shared__ double red[kPcgThreads / 32];
    __shared__ double red6[kPcgRows][kPcgThreads / 32][6];   // per (row set, warp): P^T r of the warp's 32 rows (one segment: segments are whole groups)
    __shared__ double xc_s[3 * (kCoarseMaxSeg + 1)];          // coarse solution at this chunk's nodes
    __shared__ int lc_next;                                   // next group of the chunk-local landmark pass (dynamic scheduling)
    __shared__ unsigned long long tma_bar[2];                 // mbarriers of the two per-iteration bulk copies (chain factors, index table)
    __shared__ signed char seg_of_s[kPcgRows][kPcgThreads / 32];   // segment of each (row set, warp); -1 beyond the chunk
    // the coarse operator is dropped for this solve if its Cholesky factorisation met a non-positive pivot (flag set by k_potrf_diag)
    const bool chain = w.precond != 1, coarse = w.precond == 0 && __ldcg(w.cStats + 5) == 0.0;
    const PcgSmemPlan<S> plan(d.pc_cp, d.pc_cl_max, d.pc_slots_max, chain, w.precond == 0 ? w.c_nc : 0);
    const int cps = plan.cps, Kp = plan.Kp;
    const int hseg = w.c_h, nseg = w.c_nseg;
    double* rc_s = reinterpret_cast<double*>(pcg_smem + plan.rec_off);   // coarse residual: over the factor staging area, after the chunk solve
    if (threadIdx.x < kPcgRows * (kPcgThreads / 32)) {
        const int hh = threadIdx.x / (kPcgThreads / 32), kk = threadIdx.x % (kPcgThreads / 32), r0 = 32 * kk + hh * kPcgThreads;
        seg_of_s[hh][kk] = (signed char)((r0 < d.pc_cp) ? r0 / hseg : -1);
    }
    if (threadIdx.x == 0) {
        lc_next = 0;
        mbar_init(&tma_bar[0], 1); mbar_init(&tma_bar[1], 1);
        mbar_fence_init();
    }
    unsigned tma_phase = 0;                                   // preconditioner applications so far: parity of both mbarriers
    S* vsm = reinterpret_cast<S*>(pcg_smem + plan.vec_off);
    S* rec = reinterpret_cast<S*>(pcg_smem + plan.rec_off);
    unsigned short* loc_s = reinterpret_cast<unsigned short*>(pcg_smem + plan.loc_off);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grid = gridDim.x, c = blockIdx.x, cp = d.pc_cp, gpc = cp / 32;
    const int nwarps = grid * (kPcgThreads / 32);
    const int wg = warp * grid + blockIdx.x;                 // landmark groups: consecutive groups go to different SMs
    const size_t np4 = 4 * (size_t)d.NP, nrows = (size_t)d.pc_chunks * cp;
    unsigned epoch = 0;
    double* sc = w.scal;
#ifdef BOS_PCG_TIMING
    long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif
    // ---- static per-thread data: this thread owns chunk rows tid and tid + 1024 for the whole solve ---------------------
    const int goff0 = __ldg(d.pc_goff + (size_t)c * gpc);
    const int cl0 = __ldg(d.pc_cl_ptr + c), ncl = __ldg(d.pc_cl_ptr + c + 1) - cl0;
    const int nslots = (__ldg(d.pc_goff + (size_t)(c + 1) * gpc) - goff0) * 32;
    int pose_i[kPcgRows], soff[kPcgRows], swid[kPcgRows], vx[kPcgRows];
#pragma unroll
    for (int h = 0; h < kPcgRows; h++) {
        const int r = tid + h * kPcgThreads;
        vx[h] = chain ? (r & 31) * Kp + (r >> 5) : r;              // position of the row in the shared-memory vectors
        pose_i[h] = (r < cp) ? __ldg(d.pc_row_pose + (size_t)c * cp + r) : -1;
        soff[h] = 0; swid[h] = 0;
        if (r < cp) {
            const int o = __ldg(d.pc_goff + (size_t)c * gpc + r / 32);
            soff[h] = (o - goff0) * 32 + lane;
            swid[h] = __ldg(d.pc_goff + (size_t)c * gpc + r / 32 + 1) - o;
        }
    }
    bool out_nb[kPcgRows];              // the row has a pose-pose neighbour in another chunk (pass B of the off-diagonal product)
#pragma unroll
    for (int h = 0; h < kPcgRows; h++) {
        out_nb[h] = false;
        if (pose_i[h] < 0) continue;
        const size_t R = (size_t)c * cp + tid + h * kPcgThreads;
        const int cnt = __ldg(d.pc_ncnt + R);
#pragma unroll
        for (int n = 0; n < 2; n++) {
            const int nb = __ldg(d.pc_nbr + (size_t)n * ((size_t)d.pc_chunks * cp) + R);
            if (nb >= 0 && (nb < c * cp || nb >= (c + 1) * cp)) out_nb[h] = true;
        }
        if (cnt > 2) out_nb[h] = true;
    }
    S posx[kPcgRows], posy[kPcgRows];   // the owned poses' translations: all the pose pass needs of the state, constant during the solve
#pragma unroll
    for (int h = 0; h < kPcgRows; h++) {
        posx[h] = posy[h] = S(0);
        if (pose_i[h] >= 0) load_lm<S>(d.pose, 2 * pose_i[h], posx[h], posy[h]);
    }
    auto stage_loc = [&]() {   // 16-bit landmark table indices of the chunk's edge slots (64-byte aligned runs of 32)
        const uint4* src = reinterpret_cast<const uint4*>(d.pc_loc + (size_t)goff0 * 32);
        uint4* dst = reinterpret_cast<uint4*>(loc_s);
        for (int k = tid; k < nslots / 8; k += kPcgThreads) dst[k] = __ldg(src + k);
    };
    // chain preconditioner: z = M^-1 r from the residual in shared memory into rows 9-11, then the per-row consumers: z to the
    // global buffer the next operator application gathers from, gamma = r.z and the diagonal-block part of delta = z.S z
    // FP32 views for the chain solve.  S = float: the residual rows (6-8) are the right-hand side, rows 9-11 work space and
    // result.  S = double: rows 9-11 hold two float triples, each a copy of the residual that its pass overwrites in place.
    constexpr bool kWide = sizeof(S) == 8;
    const unsigned zoff = (unsigned)(plan.vec_off + (size_t)9 * cps * sizeof(S)), roff = (unsigned)(plan.vec_off + (size_t)6 * cps * sizeof(S));
    const unsigned ch_r1 = kWide ? zoff : roff, ch_w1 = zoff, ch_r2 = kWide ? zoff + 12u * cps : roff, ch_w2 = kWide ? zoff + 12u * cps : zoff;
    const float* zres = reinterpret_cast<const float*>(pcg_smem + ch_w2);
    auto precond_chain = [&](S* zdst, double& gacc, double& dacc2) {
        __syncthreads();
        // The record / index region is free from here on (the pose pass is over): the chunk's FP32 factors (93 KB at synth-2M) arrive by 1-D TMA
        // bulk copies WHILE the coarse restriction below runs; everybody waits on the mbarrier right before the chunk solve.  A cooperative
        // copy loop (six dependent L2 round trips per thread) took 20 k cycles per iteration here.
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the region's earlier generic-proxy accesses before the async-proxy writes
            const unsigned fb = (unsigned)plan.fac_floats * 4u;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(w.chF + (size_t)c * plan.fac_floats);
            mbar_expect_tx(&tma_bar[0], fb);
            for (unsigned o = 0; o < fb; o += 16384u) tma_bulk_load(pcg_smem + plan.rec_off + o, src + o, fb - o < 16384u ? fb - o : 16384u, &tma_bar[0]);
        }
        if (coarse) {   // this chunk's part of P^T r; the exchange over the grid overlaps the chain solve below
#pragma unroll
            for (int h = 0; h < kPcgRows; h++) {
                double c6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                const int i = pose_i[h];
                if (i >= 0 && i != d.fixed) {
                    int js; float t;
                    coarse_seg(tid + h * kPcgThreads, cp, hseg, js, t);
                    const S* v = vsm + vx[h];
#pragma unroll
                    for (int a = 0; a < 3; a++) { const double rv = (double)v[(6 + a) * cps]; c6[a] = (1.0 - (double)t) * rv; c6[3 + a] = (double)t * rv; }
                }
#pragma unroll
                for (int a = 0; a < 6; a++) { const double sv = warp_sum(c6[a]); if (lane == 0) red6[h][warp][a] = sv; }
            }
            __syncthreads();
            if (tid < 6 * nseg) {   // segment js = the warps whose 32 rows lie in it
                const int js = tid / 6, a = tid - 6 * js;
                double sv = 0.0;
                for (int h = 0; h < kPcgRows; h++)
                    for (int k = 0; k < kPcgThreads / 32; k++)
                        if (seg_of_s[h][k] == js) sv += red6[h][k][a];
                __stcg(w.cRc + 6LL * ((size_t)c * nseg + js) + a, sv);
            }
            __syncthreads();
            grid_arrive(w.bar, epoch);
        }
        PCG_T(11);
        if (kWide) {
            float* c1 = reinterpret_cast<float*>(pcg_smem + ch_r1);
            float* c2 = reinterpret_cast<float*>(pcg_smem + ch_r2);
#pragma unroll
            for (int h = 0; h < kPcgRows; h++) {
                if (tid + h * kPcgThreads >= cp) continue;
                const S* v = vsm + vx[h];
#pragma unroll
                for (int a = 0; a < 3; a++) { const float f = (float)v[(6 + a) * cps]; c1[a * cps + vx[h]] = f; c2[a * cps + vx[h]] = f; }
            }
        }
        PCG_T(12);
        mbar_wait(&tma_bar[0], tma_phase & 1u);   // the factors have landed
        __syncthreads();
        PCG_T(8);
        const int nr_c = 3 * (nseg + 1);
        const bool coarse_side = coarse && nr_c <= kCoarseNr;      // the coarse solve runs beside the chain solve, on other warps
#ifndef BOS_EXP_NO_CHAIN
        if (tid < 64) chain_apply((unsigned)plan.rec_off, ch_r1, ch_w1, ch_r2, ch_w2, cps, Kp, cp / 32);
        else
#endif
        if (coarse_side && tid < 64 + kCoarseWarps * 32)
            coarse_apply(w.cAinv, w.c_ld, w.cRc, w.c_nc, grid * nseg, 3 * c * nseg, nr_c, w.bar, epoch * gridDim.x, &red6[0][0][0], xc_s);
        __syncthreads();
        PCG_T(9);
        if (coarse && !coarse_side) {
            grid_wait(w.bar, gridDim.x, epoch);
            const int nc = w.c_nc, nsegs_all = grid * nseg;      // node n = start of global segment n; the last node only closes a segment
            for (int m = tid; m < nc; m += kPcgThreads) {
                const int nd = m / 3, a = m - 3 * nd;
                double rv = 0.0;
                if (nd < nsegs_all) rv += __ldcg(w.cRc + 6LL * nd + a);
                if (nd > 0) rv += __ldcg(w.cRc + 6LL * (nd - 1) + 3 + a);
                rc_s[m] = rv;
            }
            __syncthreads();
            {   // rows of A_c^-1 for this chunk's nodes c * nseg .. (c + 1) * nseg: every warp takes a slice of one row, four loads in flight per lane
                const int nr = 3 * (nseg + 1), parts = (kPcgThreads / 32) / nr > 0 ? (kPcgThreads / 32) / nr : 1;
                const int len = ((nc + parts - 1) / parts + 31) / 32 * 32;
                if (warp < nr * parts) {
                    const int row = warp % nr, part = warp / nr;
                    const double* arow = w.cAinv + (size_t)(3 * c * nseg + row) * w.c_ld;
                    const int m1 = (part + 1) * len < nc ? (part + 1) * len : nc;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    int m = part * len + lane;
                    for (; m + 96 < m1; m += 128) {
                        const double a0 = __ldg(arow + m), a1 = __ldg(arow + m + 32), a2 = __ldg(arow + m + 64), a3 = __ldg(arow + m + 96);
                        s0 += a0 * rc_s[m]; s1 += a1 * rc_s[m + 32]; s2 += a2 * rc_s[m + 64]; s3 += a3 * rc_s[m + 96];
                    }
                    for (; m < m1; m += 32) s0 += __ldg(arow + m) * rc_s[m];
                    const double sv = warp_sum((s0 + s1) + (s2 + s3));
                    if (lane == 0) red6[0][warp][0] = sv;          // red6 is free here: reused as the per-warp partial
                }
                __syncthreads();
                if (tid < nr) {
                    double sv = 0.0;
                    for (int q = 0; q < parts; q++) sv += red6[0][q * nr + tid][0];
                    xc_s[tid] = sv;
                }
            }
            __syncthreads();
        }
        // the factors are not needed any more: the index table they overwrote comes back by a bulk copy while the per-row work below runs
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const unsigned lb = (unsigned)nslots * 2u;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(d.pc_loc + (size_t)goff0 * 32);
            mbar_expect_tx(&tma_bar[1], lb);
            for (unsigned o = 0; o < lb; o += 16384u) tma_bulk_load(reinterpret_cast<unsigned char*>(loc_s) + o, src + o, lb - o < 16384u ? lb - o : 16384u, &tma_bar[1]);
        }
        // the FP32 result sits in rows 9-11, which now become the chunk's z in S (read by the next operator application): registers first
        float zf[kPcgRows][3];
#pragma unroll
        for (int h = 0; h < kPcgRows; h++) {
            zf[h][0] = zf[h][1] = zf[h][2] = 0.f;
            if (tid + h * kPcgThreads < cp) { zf[h][0] = zres[vx[h]]; zf[h][1] = zres[cps + vx[h]]; zf[h][2] = zres[2 * cps + vx[h]]; }
        }
        __syncthreads();
        PCG_T(13);
#pragma unroll
        for (int h = 0; h < kPcgRows; h++) {
            const int i = pose_i[h];
            const int r = tid + h * kPcgThreads;
            if (r >= cp) continue;
            S* v = vsm + vx[h];
            if (i < 0) { v[9 * cps] = S(0); v[10 * cps] = S(0); v[11 * cps] = S(0); continue; }
            const S r0 = v[6 * cps], r1 = v[7 * cps], r2 = v[8 * cps];
            S zn0 = (S)zf[h][0], zn1 = (S)zf[h][1], zn2 = (S)zf[h][2];
            if (coarse && i != d.fixed) {
                int js; float tf;
                coarse_seg(r, cp, hseg, js, tf);
                const double t = (double)tf;
                const double* xl = xc_s + 3 * js;
                zn0 += (S)((1.0 - t) * xl[0] + t * xl[3]); zn1 += (S)((1.0 - t) * xl[1] + t * xl[4]); zn2 += (S)((1.0 - t) * xl[2] + t * xl[5]);
            }
            const S* hp = w.rowS + (size_t)c * cp + r;
            const S h0 = __ldg(hp), h1 = __ldg(hp + nrows), h2 = __ldg(hp + 2 * nrows), h3 = __ldg(hp + 3 * nrows), h4 = __ldg(hp + 4 * nrows),
                    h5 = __ldg(hp + 5 * nrows);
            st4cg(zdst + 4LL * i, zn0, zn1, zn2);
            v[9 * cps] = zn0; v[10 * cps] = zn1; v[11 * cps] = zn2;
            gacc += (double)r0 * (double)zn0 + (double)r1 * (double)zn1 + (double)r2 * (double)zn2;
            dacc2 += (double)zn0 * (double)(h0 * zn0 + h1 * zn1 + h2 * zn2) + (double)zn1 * (double)(h1 * zn0 + h3 * zn1 + h4 * zn2) +
                     (double)zn2 * (double)(h2 * zn0 + h4 * zn1 + h5 * zn2);
        }
        PCG_T(14);
        mbar_wait(&tma_bar[1], tma_phase & 1u);   // the index table is back
        tma_phase++;
        PCG_T(10);
    };
    {
        stage_loc();
#pragma unroll
        for (int h = 0; h < kPcgRows; h++) {
            const int r = tid + h * kPcgThreads;
            if (r < cp) {
#pragma unroll
                for (int k = 0; k < 12; k++) vsm[(size_t)k * cps + vx[h]] = (pose_i[h] >= 0 && k >= 6 && k < 9) ? w.rS[(size_t)(k - 6) * nrows + (size_t)c * cp + r] : S(0);
                if (!chain && pose_i[h] >= 0) {   // z_0 = M^-1 g of the 3x3 flavour comes from k_pcg_fused_prep; the chain flavour computes its own below
                    S a0, a1, a2, ap;
                    ld4cg(w.z4 + 4LL * pose_i[h], a0, a1, a2, ap);
                    vsm[(size_t)9 * cps + vx[h]] = a0; vsm[(size_t)10 * cps + vx[h]] = a1; vsm[(size_t)11 * cps + vx[h]] = a2;
                }
            }
        }
    }
    if (chain) {   // z_0 = M^-1 g, gamma_0 and the first part of delta_0 (the 3x3 flavour gets them from k_pcg_fused_prep)
        double g0 = 0.0, d0 = 0.0;
        precond_chain(w.z4, g0, d0);
        const double sg = block_sum_pcg(g0, red);
        const double sd = block_sum_pcg(d0, red);
        if (tid == 0) {
            if (sg != 0.0) atomicAdd(sc + FS_GAMMA0, sg);
            if (sd != 0.0) atomicAdd(sc + FS_DELTA0, sd);
        }
        grid_barrier(w.bar, gridDim.x, epoch);
    }
    __syncthreads();
    const double gamma_init = __ldcg(sc + FS_GAMMA0);
    double gamma_prev = 1.0, alpha_prev = 1.0;
    int it = 0;
    bool bad = false;
    S* pxy = rec;                                   // [cp][2] the chunk's pose positions, in the record region while the landmark pass runs
    if (gamma_init > 0.0) {
        for (; it < max_iters;) {
            const int cur = it % 3, nxt = (it + 1) % 3, nn = (it + 2) % 3;
            const S* zc = w.z4 + (size_t)(it & 1) * np4;
            S* zn = w.z4 + (size_t)((it + 1) & 1) * np4;
            S* tp = w.tpart + (size_t)(it & 1) * 2 * (size_t)d.n_q;
            // ---- phase L: the chunk's partial t_l of every landmark it sees (shared memory only), off-diagonal pose-pose products -----------
#pragma unroll
            for (int h = 0; h < kPcgRows; h++) {
                const int r = tid + h * kPcgThreads;
                if (r < cp) { pxy[2 * r] = posx[h]; pxy[2 * r + 1] = posy[h]; }
            }
            __syncthreads();
            double dacc = 0.0;
            S yo[kPcgRows][3];       // off-diagonal pose-pose product of the owned rows, in registers until the pose pass
            // Pass A (before the grid barrier): neighbours inside this chunk, whose z is in shared memory (rows 9-11).  Neighbours in other chunks
            // (the two ends of the chunk, loop closures) follow in pass B, after the barrier that makes every chunk's z visible: the end-of-iteration
            // barrier of round 1 is gone.  The row's static data (neighbour ids, both inline blocks) is fetched in ONE batch of independent loads:
            // rows without a neighbour hold zero blocks.
            auto off_row = [&](int h, bool outside, S& y0, S& y1, S& y2) {
                const int i = pose_i[h];
                const int r = tid + h * kPcgThreads;
                const size_t R = (size_t)c * cp + r;
                const int cnt = __ldg(d.pc_ncnt + R);
                const int nbv[2] = {__ldg(d.pc_nbr + R), __ldg(d.pc_nbr + nrows + R)};
                S bb[2][6];
#pragma unroll
                for (int n = 0; n < 2; n++) {
                    const S* o = w.rowS + (size_t)(12 + 6 * n) * nrows + R;
#pragma unroll
                    for (int k = 0; k < 6; k++) bb[n][k] = __ldg(o + (size_t)k * nrows);
                }
                auto nbr_z = [&](int nb, S& n0, S& n1, S& n2) -> bool {      // false: the neighbour belongs to the other pass
                    const int rr = nb - c * cp;
                    const bool inside = rr >= 0 && rr < cp;
                    if (inside == outside) return false;
                    if (inside) {
                        const int xv = chain ? (rr & 31) * Kp + (rr >> 5) : rr;
                        n0 = vsm[(size_t)9 * cps + xv]; n1 = vsm[(size_t)10 * cps + xv]; n2 = vsm[(size_t)11 * cps + xv];
                    } else {
                        S np_;
                        ld4cg(zc + 4LL * nb, n0, n1, n2, np_);
                    }
                    return true;
                };
#ifndef BOS_EXP_NO_OFFDIAG
#pragma unroll
                for (int n = 0; n < 2; n++) {
                    const int nb = nbv[n];
                    S n0, n1, n2;
                    if (nb < 0 || !nbr_z(nb, n0, n1, n2)) continue;
                    y0 += bb[n][0] * n0 + bb[n][1] * n1 + bb[n][2] * n2;
                    y1 += bb[n][1] * n0 + bb[n][3] * n1 + bb[n][4] * n2;
                    y2 += bb[n][2] * n0 + bb[n][4] * n1 + bb[n][5] * n2;
                }
                if (cnt > 2) {   // loop closures beyond the two inline neighbours: generic adjacency
                    const int q0 = __ldg(d.pp_ptr + i);
                    for (int q = q0 + 2; q < q0 + cnt; q++) {
                        const int nb = __ldg(d.pp_nbr + q);
                        const int sl = __ldg(d.pp_slot + q);
                        const S* Bo = d.Hoff + 9LL * (sl & 0x7fffffff);
                        S n0, n1, n2;
                        if (!nbr_z(nb, n0, n1, n2)) continue;
                        if (sl >= 0) {
                            y0 += Bo[0] * n0 + Bo[1] * n1 + Bo[2] * n2;
                            y1 += Bo[3] * n0 + Bo[4] * n1 + Bo[5] * n2;
                            y2 += Bo[6] * n0 + Bo[7] * n1 + Bo[8] * n2;
                        } else {
                            y0 += Bo[0] * n0 + Bo[3] * n1 + Bo[6] * n2;
                            y1 += Bo[1] * n0 + Bo[4] * n1 + Bo[7] * n2;
                            y2 += Bo[2] * n0 + Bo[5] * n1 + Bo[8] * n2;
                        }
                    }
                }
#endif
            };
#pragma unroll
            for (int h = 0; h < kPcgRows; h++) {
                yo[h][0] = yo[h][1] = yo[h][2] = S(0);
                if (pose_i[h] < 0) continue;
                S y0 = S(0), y1 = S(0), y2 = S(0);
                off_row(h, false, y0, y1, y2);
                const S* v = vsm + vx[h];
                dacc += (double)v[9 * cps] * (double)y0 + (double)v[10 * cps] * (double)y1 + (double)v[11 * cps] * (double)y2;
                yo[h][0] = y0; yo[h][1] = y1; yo[h][2] = y2;
            }
            PCG_T(0);
#ifndef BOS_EXP_NO_LROWS
            pcg_local_landmark_rows<S>(d, w, c, vsm + (size_t)9 * cps, cps, Kp, chain, pxy, tp, &lc_next);
#endif
            PCG_T(1);
            grid_barrier(w.bar, gridDim.x, epoch);       // every chunk's partials, and the z and gamma the previous iteration left, are visible
            if (tid == 0) lc_next = 0;                   // every warp is past the landmark pass; the next one is several CTA barriers away
            PCG_T(3);
            if (it > 0 && !(__ldcg(sc + FS_GAMMA0 + cur) > tol2 * gamma_init)) break;      // converged: x is final (the L phase above was for nothing)
            // pass B of the off-diagonal product: neighbours in other chunks
#pragma unroll
            for (int h = 0; h < kPcgRows; h++) {
                if (!out_nb[h]) continue;
                S y0 = S(0), y1 = S(0), y2 = S(0);
                off_row(h, true, y0, y1, y2);
                const S* v = vsm + vx[h];
                dacc += (double)v[9 * cps] * (double)y0 + (double)v[10 * cps] * (double)y1 + (double)v[11 * cps] * (double)y2;
                yo[h][0] += y0; yo[h][1] += y1; yo[h][2] += y2;
            }
            // ---- phase P, part 1: stage the chunk's landmark records {u_l, position}: t_l = the partials of all chunks that see the landmark ----
            {
                double tu = 0.0;
                const size_t nq = (size_t)d.n_q;
                for (int k = tid; k < ncl; k += kPcgThreads) {
                    const int q = cl0 + k;
                    int src[kShareEll];
#pragma unroll
                    for (int j = 0; j < kShareEll; j++) src[j] = __ldg(d.sh_ell + (size_t)j * nq + q);      // one round trip: all sources
                    const S i00 = __ldg(w.qstat + q), i01 = __ldg(w.qstat + nq + q), i11 = __ldg(w.qstat + 2 * nq + q);
                    const S lx = __ldg(w.qstat + 3 * nq + q), ly = __ldg(w.qstat + 4 * nq + q);
                    S t0 = S(0), t1 = S(0);
#pragma unroll
                    for (int j = 0; j < kShareEll; j++)
                        if (src[j] >= 0) {
                            const typename Vec2T<S>::type v2 = __ldcg(reinterpret_cast<const typename Vec2T<S>::type*>(tp) + src[j]);
                            t0 += v2.x; t1 += v2.y;
                        }
                    if (src[kShareEll - 1] >= 0)   // a landmark seen from more than kShareEll chunks: the rest of its list
                        for (int qq = __ldg(d.sh_ptr + q) + kShareEll; qq < __ldg(d.sh_ptr + q + 1); qq++) {
                            const typename Vec2T<S>::type v2 = __ldcg(reinterpret_cast<const typename Vec2T<S>::type*>(tp) + __ldg(d.sh_src + qq));
                            t0 += v2.x; t1 += v2.y;
                        }
                    const S u0 = i00 * t0 + i01 * t1, u1 = i01 * t0 + i11 * t1;
                    rec[4 * k] = u0; rec[4 * k + 1] = u1; rec[4 * k + 2] = lx; rec[4 * k + 3] = ly;
                    if (d.sh_first[q]) tu += (double)t0 * (double)u0 + (double)t1 * (double)u1;     // counted once per landmark
                }
                // delta = z . S z is complete once every chunk has added its  z . (off-diagonal part) - t . u  (the diagonal part came with z)
                const double sdel = block_sum_pcg(dacc - tu, red);
                if (tid == 0) {
                    if (sdel != 0.0) atomicAdd(sc + FS_DELTA0 + cur, sdel);
                    __threadfence();
                }
                __syncthreads();                           // rec is staged; thread 0's atomic is ordered before the arrive
                grid_arrive(w.bar, epoch);
            }
            PCG_T(7);
            // ---- part 2: w = S z for the owned rows while the delta exchange is in flight -------------------------------------------
            const S so_u = (S)w.sqrt_omega;
            S wv[kPcgRows][3];
#pragma unroll
            for (int h = 0; h < kPcgRows; h++) {
                const int i = pose_i[h];
                const int r = tid + h * kPcgThreads;
                wv[h][0] = wv[h][1] = wv[h][2] = S(0);
                if (r >= cp) continue;                       // warp-uniform: cp is a multiple of 32
                const int W = swid[h];
                S w0 = S(0), w1 = S(0), w2 = S(0);
                const bool active = i >= 0 && i != d.fixed;
                for (int t = 0; t < W; t++) {
                    const unsigned lc = loc_s[soff[h] + t * 32];
                    if (lc == 0xffffu || !active) continue;
                    const S u0 = rec[4 * lc], u1 = rec[4 * lc + 1], lx = rec[4 * lc + 2], ly = rec[4 * lc + 3];
                    S j0, j1;
                    bearing_jl_world<S>(posx[h], posy[h], lx, ly, j0, j1);
                    const S so = w.omega_uniform ? so_u : __ldg(w.Pw + (size_t)goff0 * 32 + soff[h] + t * 32);
                    j0 *= so; j1 *= so;
                    const S m = j0 * u0 + j1 * u1;
                    w0 += j0 * m; w1 += j1 * m; w2 -= (j0 * ly - j1 * lx) * m;
                }
                if (i < 0) continue;
                const S* hp = w.rowS + (size_t)c * cp + r;
                const S h0 = __ldg(hp), h1 = __ldg(hp + nrows), h2 = __ldg(hp + 2 * nrows), h3 = __ldg(hp + 3 * nrows), h4 = __ldg(hp + 4 * nrows),
                        h5 = __ldg(hp + 5 * nrows);
                const S* v = vsm + vx[h];
                const S z0 = v[9 * cps], z1 = v[10 * cps], z2 = v[11 * cps];
                wv[h][0] = w0 + h0 * z0 + h1 * z1 + h2 * z2 + yo[h][0];
                wv[h][1] = w1 + h1 * z0 + h3 * z1 + h4 * z2 + yo[h][1];
                wv[h][2] = w2 + h2 * z0 + h4 * z1 + h5 * z2 + yo[h][2];
            }
            PCG_T(4);
            grid_wait(w.bar, gridDim.x, epoch);
            const double gamma = __ldcg(sc + FS_GAMMA0 + cur), delta = __ldcg(sc + FS_DELTA0 + cur);
            const double beta = (it == 0) ? 0.0 : gamma / gamma_prev;
            const double denom = (it == 0) ? delta : delta - beta * gamma / alpha_prev;
            if (!(denom > 0.0)) { bad = true; break; }
            const double alpha = gamma / denom;
            if (blockIdx.x == 0 && tid == 0) { __stcg(sc + FS_GAMMA0 + nn, 0.0); __stcg(sc + FS_DELTA0 + nn, 0.0); }
            const S al = (S)alpha, be = (S)beta;
            double gacc = 0.0, dacc2 = 0.0;
            // ---- part 3: the recurrences of the owned rows ---------------------------------------------------------------------------
#pragma unroll
            for (int h = 0; h < kPcgRows; h++) {
                const int i = pose_i[h];
                const int r = tid + h * kPcgThreads;
                if (r >= cp || i < 0) continue;
                S* v = vsm + vx[h];
                const S z0 = v[9 * cps], z1 = v[10 * cps], z2 = v[11 * cps];
                const S p0 = z0 + be * v[0], p1 = z1 + be * v[cps], p2 = z2 + be * v[2 * cps];
                const S s0_ = wv[h][0] + be * v[3 * cps], s1 = wv[h][1] + be * v[4 * cps], s2 = wv[h][2] + be * v[5 * cps];
                v[0] = p0; v[cps] = p1; v[2 * cps] = p2;
                v[3 * cps] = s0_; v[4 * cps] = s1; v[5 * cps] = s2;
                const S r0 = v[6 * cps] - al * s0_, r1 = v[7 * cps] - al * s1, r2 = v[8 * cps] - al * s2;
                v[6 * cps] = r0; v[7 * cps] = r1; v[8 * cps] = r2;
                S* xg = w.xS + (size_t)c * cp + r;
                xg[0] += al * p0; xg[nrows] += al * p1; xg[2 * nrows] += al * p2;
                if (chain) continue;                         // z = M^-1 r needs the whole chunk's residual: below
                const S* hp = w.rowS + (size_t)c * cp + r;
                const S h0 = __ldg(hp), h1 = __ldg(hp + nrows), h2 = __ldg(hp + 2 * nrows), h3 = __ldg(hp + 3 * nrows), h4 = __ldg(hp + 4 * nrows),
                        h5 = __ldg(hp + 5 * nrows);
                const S* mi = hp + 6 * nrows;
                const S m0 = __ldg(mi), m1 = __ldg(mi + nrows), m2 = __ldg(mi + 2 * nrows), m3 = __ldg(mi + 3 * nrows), m4 = __ldg(mi + 4 * nrows),
                        m5 = __ldg(mi + 5 * nrows);
                const S zn0 = m0 * r0 + m1 * r1 + m2 * r2, zn1 = m1 * r0 + m3 * r1 + m4 * r2, zn2 = m2 * r0 + m4 * r1 + m5 * r2;
                st4cg(zn + 4LL * i, zn0, zn1, zn2);
                v[9 * cps] = zn0; v[10 * cps] = zn1; v[11 * cps] = zn2;      // the chunk's copy for the next operator application
                gacc += (double)r0 * (double)zn0 + (double)r1 * (double)zn1 + (double)r2 * (double)zn2;
                dacc2 += (double)zn0 * (double)(h0 * zn0 + h1 * zn1 + h2 * zn2) + (double)zn1 * (double)(h1 * zn0 + h3 * zn1 + h4 * zn2) +
                         (double)zn2 * (double)(h2 * zn0 + h4 * zn1 + h5 * zn2);
            }
            if (chain) precond_chain(zn, gacc, dacc2);
            {
                const double sg = block_sum_pcg(gacc, red);
                const double sd2 = block_sum_pcg(dacc2, red);
                if (tid == 0) {
                    if (sg != 0.0) atomicAdd(sc + FS_GAMMA0 + nxt, sg);
                    if (sd2 != 0.0) atomicAdd(sc + FS_DELTA0 + nxt, sd2);
                }
            }
            PCG_T(5);
            // no barrier here: the next iteration's first grid barrier orders this iteration's z, gamma and delta parts before their readers
            gamma_prev = gamma; alpha_prev = alpha;
            it++;
        }
    }
    // ---- epilogue: dx_p = x; x as padded records for the back-substitution gather; dx_l ----------------------------------
    S* x4 = w.z4 + (size_t)((it + 1) & 1) * np4;    // the z buffer that is not current
#pragma unroll
    for (int h = 0; h < kPcgRows; h++) {
        const int i = pose_i[h];
        if (i < 0) continue;
        const S* xg = w.xS + (size_t)c * cp + tid + h * kPcgThreads;
        const S x0 = xg[0], x1 = xg[nrows], x2 = xg[2 * nrows];
        d.delta[3LL * i] = x0; d.delta[3LL * i + 1] = x1; d.delta[3LL * i + 2] = x2;
        st4cg(x4 + 4LL * i, x0, x1, x2);
    }
    const int gtid = blockIdx.x * kPcgThreads + tid, gsz = gridDim.x * kPcgThreads;
    for (int l = gtid; l < d.NL; l += gsz)
        if (__ldg(d.tri_ptr + l) == __ldg(d.tri_ptr + l + 1)) {   // unobserved landmark: dx_l = -Hll^-1 b_l
            d.delta[3LL * d.NP + 2LL * l] = -w.ul[2LL * l];
            d.delta[3LL * d.NP + 2LL * l + 1] = -w.ul[2LL * l + 1];
        }
    grid_barrier(w.bar, gridDim.x, epoch);
    double unused = 0.0;
    pcg_landmark_rows<S, 1>(d, w, x4, wg, nwarps, unused);
    if (gtid == 0) { sc[SC_ITER] = (double)it; sc[SC_BAD] = bad ? 1.0 : 0.0; }
#ifdef BOS_PCG_TIMING
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1 || blockIdx.x == gridDim.x / 2))
        printf("cta %d iters %d cycles/iter: offdiag %lld Lrows %lld bsum %lld bar1 %lld stage %lld rows %lld bsum2 %lld bar2 %lld | chain: restrict %lld f32copy %lld tma-wait %lld solve %lld | finish: zload %lld rows %lld loc-wait %lld\n",
               (int)blockIdx.x, it, tacc[0] / (it ? it : 1), tacc[1] / (it ? it : 1), tacc[2] / (it ? it : 1), tacc[3] / (it ? it : 1), tacc[7] / (it ? it : 1),
               tacc[4] / (it ? it : 1), tacc[5] / (it ? it : 1), tacc[6] / (it ? it : 1), tacc[11] / (it ? it : 1), tacc[12] / (it ? it : 1), tacc[8] / (it ? it : 1),
               tacc[9] / (it ? it : 1), tacc[13] / (it ? it : 1), tacc[14] / (it ? it : 1), tacc[10] / (it ? it : 1));
#endif
}

// the persistent kernel needs the chunk's vectors, landmark records and slot indices in shared memory, at most two rows per thread
template <typename S>
bool pcg_fused_supported(const Dev<S>& d) {
    if (!d.pc_ok || d.pc_cp > kPcgRows * kPcgThreads || d.pc_chunks < 1) return false;
    const PcgSmemPlan<S> plan(d.pc_cp, d.pc_cl_max, d.pc_slots_max, false);
    return plan.bytes <= (size_t)kPcgSmemBudget;
}
template <typename S>
bool pcg_chain_supported(const Dev<S>& d, int nc_coarse) {
    const PcgSmemPlan<S> plan(d.pc_cp, d.pc_cl_max, d.pc_slots_max, true, nc_coarse);
    return plan.bytes <= (size_t)kPcgSmemBudget && d.pc_cp / 32 <= 64;
}

template <typename S>
int launch_pcg_fused(const Dev<S>& d, PcgWork<S>& w, int max_iters, double rtol, cudaStream_t st, int* iterations_out, int* launches) {
    int nl = 0;
    const int gl = (d.NL + 255) / 256;
    const long long nrows = (long long)d.pc_chunks * d.pc_cp;
    cudaMemsetAsync(w.scal, 0, 32 * sizeof(double), st);
    cudaMemsetAsync(w.bar, 0, 4 * sizeof(unsigned), st);
    cudaMemsetAsync(w.xS, 0, 3 * (size_t)nrows * sizeof(S), st);
    if (d.NL > 0) { k_lm_prep<S><<<gl, 256, 0, st>>>(d, w.hllinv, w.ul); nl++; }
    if (d.n_clm > 0) { k_ell_fill<S><<<(std::max(d.n_clm, d.n_q) + 255) / 256, 256, 0, st>>>(d, w); nl++; }
    // FP32 flavour: the Schur diagonal blocks are differences of terms ~1e8 times larger than their small eigenvalues (world-frame
    // lever arms); in float they are not reliably positive definite, and a block-tridiagonal factorisation built on them breaks
    // down at synth-2M.  The 3x3 block-Jacobi preconditioner (which only inverts them) is what the FP32 path runs.
    int precond = (w.precond != 1 && sizeof(S) == 8 && pcg_chain_supported<S>(d, 0)) ? w.precond : 1;
    if (precond == 0 && !pcg_chain_supported<S>(d, w.c_nc)) precond = 2;
    if (precond == 0 && (w.c_nseg < 1 || w.c_nseg > kCoarseMaxSeg || (w.c_bw <= 0 && w.c_nc > 3 * 160))) precond = 2;
    const int precond_asked = w.precond;
    w.precond = precond;
    w.precond_used = precond;
    k_pcg_fused_prep<S><<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(d, w); nl++;
    if (precond != 1) { k_pcg_chain_factor<S><<<d.pc_chunks, 128, 0, st>>>(d, w); nl++; }
    // coarse operator A_c = P^T S P, its Cholesky factor and explicit inverse.  Any SPD coarse operator makes a valid preconditioner, so A_c^-1 is
    // kept across GN steps.  It is rebuilt (a) after the state was replaced from outside or a breakdown (coarse_valid), (b) when the CG iteration
    // count drifts 25 % above what it was right after the last rebuild (coarse_stale), (c) for every solve with pcg_coarse_refresh = 1, and
    // otherwise (d) when it PAYS: a rebuild costs coarse_ratio CG iterations (measured below: 2.65 ms against 86 us per iteration at synth-2M = 31),
    // a stale operator costs its - its_ref extra iterations per solve; rebuild once the extra iterations spent since the last rebuild add up to
    // the price of one (the break-even rule of rent-or-buy), at the latest after 8 x pcg_coarse_refresh solves.
    const bool every = w.coarse_refresh <= 1;
    const bool due = every || w.coarse_age >= 8 * w.coarse_refresh || (w.coarse_age >= 2 && w.coarse_excess >= w.coarse_ratio);
    const bool refresh = precond == 0 && (!w.coarse_valid || w.coarse_stale || due);
    bool timed = false;
    if (refresh) {
        timed = true;
        for (int k = 0; k < 3 && timed; k++)
            if (!w.coarse_ev[k] && cudaEventCreate(&w.coarse_ev[k]) != cudaSuccess) { w.coarse_ev[k] = nullptr; cudaGetLastError(); timed = false; }
        if (timed) cudaEventRecord(w.coarse_ev[0], st);
    }
    if (refresh) {
        const int nc = w.c_nc;
        cudaMemsetAsync(w.cA, 0, sizeof(double) * (size_t)nc * nc, st);
        cudaMemsetAsync(w.cStats, 0, sizeof(double) * 8, st);
        if (d.n_clm > 0) { k_coarse_lm<S><<<(d.n_clm + 127) / 128, 128, 0, st>>>(d, w); nl++; }
        k_coarse_pose<S><<<d.pc_chunks * w.c_nseg, 256, 0, st>>>(d, w); nl++;
        k_coarse_fix<<<(nc + 127) / 128, 128, 0, st>>>(w.cA, nc); nl++;
        if (w.c_bw > 0) {   // band matrix: one CTA factorises it, one warp per column inverts it
            const int W = w.c_bw + 1;
            const size_t csm = sizeof(double) * (size_t)(W + 1) * W;
            if (!ensure_dyn_smem((const void*)k_coarse_band_chol, csm)) return -1;
            k_coarse_band_chol<<<1, kBandCholThreads, csm, st>>>(w.cA, nc, w.c_bw, w.cLc, w.cLr, w.cLdi, w.cStats); nl++;
            constexpr int IW = 4;
            const size_t ism = sizeof(double) * IW * (size_t)nc;
            if (!ensure_dyn_smem((const void*)k_coarse_band_inverse<IW>, ism)) return -1;
            k_coarse_band_inverse<IW><<<(nc + IW - 1) / IW, IW * 32, ism, st>>>(w.cLc, w.cLr, w.cLdi, w.cAinv, nc, w.c_bw, w.c_ld); nl++;
        } else {            // wide coupling (loop closures far along the chain): dense Cholesky + dense triangular inverse, nc <= 480
            nl += dense_cholesky_lower<double>(w.cA, nc, w.cStats, st);
            constexpr int IW = 4;
            const size_t ism = sizeof(double) * IW * (size_t)nc;
            k_coarse_inverse<IW><<<(nc + IW - 1) / IW, IW * 32, ism, st>>>(w.cA, w.cAinv, nc, w.c_ld); nl++;
        }
        w.coarse_valid = true; w.coarse_age = 0; w.coarse_stale = false;
        if (timed) cudaEventRecord(w.coarse_ev[1], st);
    }
    if (precond == 0) w.coarse_age++;
    const PcgSmemPlan<S> plan(d.pc_cp, d.pc_cl_max, d.pc_slots_max, precond != 1, precond == 0 ? w.c_nc : 0);
    if (plan.bytes > (size_t)kPcgSmemBudget) return -1;
    if (!ensure_dyn_smem((const void*)k_pcg_fused<S>, plan.bytes)) return -1;
    Dev<S> dd = d;
    PcgWork<S> ww = w;
    w.precond = precond_asked;
    double tol2 = rtol * rtol;
    void* args[] = {(void*)&dd, (void*)&ww, (void*)&max_iters, (void*)&tol2};
    if (cudaLaunchCooperativeKernel((const void*)k_pcg_fused<S>, dim3(d.pc_chunks), dim3(kPcgThreads), args, plan.bytes, st) != cudaSuccess) return -1;
    nl++;
    if (timed) cudaEventRecord(w.coarse_ev[2], st);
    double host_pageable[32];
    double* host_scal = w.host_scal ? w.host_scal : host_pageable;   // pinned: a true asynchronous copy instead of a staged one
    if (cudaMemcpyAsync(host_scal, w.scal, 32 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if (iterations_out) *iterations_out = (int)host_scal[SC_ITER];
    if (launches) *launches = nl;
    if (precond == 0) {   // iteration count right after a rebuild is the yardstick for the following, lagged solves
        const int its = (int)host_scal[SC_ITER];
        if (refresh) {
            w.coarse_its_ref = its;
            w.coarse_excess = 0.0;
            float t_refresh = 0.f, t_solve = 0.f;
            if (timed && its > 0 && cudaEventElapsedTime(&t_refresh, w.coarse_ev[0], w.coarse_ev[1]) == cudaSuccess &&
                cudaEventElapsedTime(&t_solve, w.coarse_ev[1], w.coarse_ev[2]) == cudaSuccess && t_solve > 0.f)
                w.coarse_ratio = std::min(256.0, std::max(4.0, (double)t_refresh / ((double)t_solve / its)));
        } else {
            if (its > w.coarse_its_ref) w.coarse_excess += its - w.coarse_its_ref;
            if (its > w.coarse_its_ref + w.coarse_its_ref / 4 + 4) w.coarse_stale = true;
        }
        w.coarse_its_last = its;
        if (host_scal[SC_BAD] != 0.0) w.coarse_valid = false;
    }
    return host_scal[SC_BAD] != 0.0 ? 1 : 0;
}

template <typename S>
int launch_pcg_solve(const Dev<S>& d, PcgWork<S>& w, int max_iters, double rtol, cudaStream_t st,
                     int* iterations_out, int* launches) {
    w.resolves = 0;
    if (w.variant == 0 && pcg_fused_supported<S>(d)) {
        int rc = launch_pcg_fused<S>(d, w, max_iters, rtol, st, iterations_out, launches);
        if (rc == 1 && w.precond != 1) {
            // breakdown (a CG denominator was not positive) under the chain / coarse preconditioner: solve again with the 3x3 blocks,
            // whose only numerical ingredient is the inverse of each Schur diagonal block
            const int asked = w.precond;
            int nl2 = 0, it2 = 0;
            w.precond = 1;
            w.resolves = 1;
            rc = launch_pcg_fused<S>(d, w, max_iters, rtol, st, &it2, &nl2);
            w.precond = asked;
            if (iterations_out) *iterations_out += it2;
            if (launches) *launches += nl2;
        }
        return rc;
    }
    int nl = 0;
    w.precond_used = 1;
    const int n = 3 * d.NP;
    const int gp = (d.NP + 255) / 256, gl = (d.NL + 255) / 256, gh = (d.n_hpl + 255) / 256, gn = (n + 255) / 256;
    cudaMemsetAsync(w.scal, 0, 16 * sizeof(double), st);
    if (d.NL > 0) { k_lm_prep<S><<<gl, 256, 0, st>>>(d, w.hllinv, w.ul); nl++; }
    if (d.n_hpl > 0) {
        k_copy_hlp<S><<<gh, 256, 0, st>>>(d.n_hpl, d.hpl_ld, d.Hpl, d.lm_order, w.Hlp);
        nl++;
    }
    k_pcg_pose_prep<S><<<gp, 256, 0, st>>>(d, w); nl++;
    k_pcg_begin<<<1, 1, 0, st>>>(w.scal, rtol); nl++;
    int dot_grid = gn < 592 ? gn : 592;
    double host_scal[16];
    const int check_every = 16;
    int it = 0;
    while (it < max_iters) {
        int chunk = max_iters - it < check_every ? max_iters - it : check_every;
        for (int c = 0; c < chunk; c++, it++) {
            const int parity = it & 1;
            k_pcg_y_init<S><<<gp, 256, 0, st>>>(d, w, parity);
            if (d.n_hpl > 0) {
                k_lm_gather<S><<<gh, 256, 0, st>>>(d.n_hpl, w.Hlp, d.hpl_ld, nullptr, d.lm_order_pose, d.lm_order_lm, w.p0, w.tl, w.scal + SC_DONE);
                k_pcg_pose_scatter<S><<<gh, 256, 0, st>>>(d, w);
                nl += 2;
            }
            k_pcg_dot<S><<<dot_grid, 256, 0, st>>>(n, w);
            k_pcg_update<S><<<gp, 256, 0, st>>>(d, w, parity);
            k_pcg_dir<S><<<gn, 256, 0, st>>>(n, w, parity);
            k_pcg_promote<<<1, 1, 0, st>>>(w.scal);
            nl += 5;
        }
        if (cudaMemcpyAsync(host_scal, w.scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
        if (host_scal[SC_DONE] != 0.0) break;
    }
    if (it == 0) {
        cudaMemcpyAsync(host_scal, w.scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    }
    if (iterations_out) *iterations_out = (int)host_scal[SC_ITER];
    // dx_p = x ; dx_l by back-substitution
    cudaMemcpyAsync(d.delta, w.x, sizeof(S) * (size_t)n, cudaMemcpyDeviceToDevice, st);
    if (d.NL > 0) {
        cudaMemsetAsync(w.tl, 0, sizeof(S) * 2 * (size_t)d.NL, st);
        if (d.n_hpl > 0) {
            k_lm_gather<S><<<gh, 256, 0, st>>>(d.n_hpl, w.Hlp, d.hpl_ld, nullptr, d.lm_order_pose, d.lm_order_lm, w.x, w.tl, nullptr);
            nl++;
        }
        k_lm_backsub<S><<<gl, 256, 0, st>>>(d, w.hllinv, w.tl); nl++;
    }
    if (launches) *launches = nl;
    return host_scal[SC_BAD] != 0.0 ? 1 : 0;
}

template int launch_pcg_solve<double>(const Dev<double>&, PcgWork<double>&, int, double, cudaStream_t, int*, int*);
template int launch_pcg_solve<float>(const Dev<float>&, PcgWork<float>&, int, double, cudaStream_t, int*, int*);

}  // namespace bos
