// solve_dense.cu -- K4/K5a/K6 for small problems: the 2x2 landmark blocks are Schur-complemented out
// EXPLICITLY into a dense pose system, which is Cholesky-factorised and solved; landmarks follow by
// back-substitution.  Replaces SimplicialLDLT on H_nofixed (slam/solver.hpp:72, solver.cpp:77-94):
// exact block elimination + exact factorisation give the same dx up to rounding.
//
// The factorisation is a right-looking blocked Cholesky (panel 64): diagonal-block POTRF in one CTA,
// row-parallel TRSM on the panel, and the trailing update C -= P P^T as 64x64 tiles -- the one real dense
// contraction on the path, run on the FP64 tensor pipe (mma.sync m8n8k4 f64 = DMMA; tcgen05 has no FP64
// kind).  S is column-major, lower triangle only.
#include "bos_internal.h"

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "bos_math.cuh"
#include "bos_schur.cuh"

namespace bos {

constexpr int NB = kDenseNB;  // panel width / tile edge (64)
constexpr int LDT = NB + 4;   // smem leading dimension: == 4 (mod 16) doubles -> conflict-free DMMA fragment loads

// ---- assembly of the dense reduced system -----------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(256) k_dense_fill_diag(Dev<S> d, S* __restrict__ Sm, S* __restrict__ g, int ld) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.NP) return;
    const S* h = d.Hpp + 6LL * p;
    const S m[9] = {h[0], h[1], h[2], h[1], h[3], h[4], h[2], h[4], h[5]};
    for (int a = 0; a < 3; a++) {
        for (int c = 0; c < 3; c++) if (a >= c) Sm[(size_t)(3 * p + a) + (size_t)(3 * p + c) * ld] = m[a * 3 + c];   // lower triangle only
        g[3 * p + a] = -d.b[3LL * p + a];
    }
}
template <typename S>
__global__ void __launch_bounds__(256) k_dense_fill_off(Dev<S> d, S* __restrict__ Sm, int ld) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.n_off) return;
    const int lo = d.off_lo[k], hi = d.off_hi[k];
    const S* B = d.Hoff + 9LL * k;  // H[lo][hi]; the lower triangle holds its transpose at (hi, lo)
    for (int a = 0; a < 3; a++)
        for (int c = 0; c < 3; c++) Sm[(size_t)(3 * hi + c) + (size_t)(3 * lo + a) * ld] = B[a * 3 + c];
}

// one warp per landmark: S[pi,pj] -= Hpl_i Hll^-1 Hpl_j^T for every pair of its observing poses (pi >= pj),
// g[pi] += Hpl_i Hll^-1 b_l
template <typename S>
__global__ void __launch_bounds__(256) k_dense_schur(Dev<S> d, const S* __restrict__ hllinv, const S* __restrict__ ul,
                                                     S* __restrict__ Sm, S* __restrict__ g, int ldS) {
    const int lane = threadIdx.x & 31;
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (l >= d.NL) return;
    const int k0 = d.lm_ptr[l], m = d.lm_ptr[l + 1] - k0;
    const S i00 = hllinv[3LL * l], i01 = hllinv[3LL * l + 1], i11 = hllinv[3LL * l + 2];
    const S u0 = ul[2LL * l], u1 = ul[2LL * l + 1];
    for (int i = lane; i < m; i += 32) {
        const S* B = d.Hpl + d.lm_order[k0 + i];
        const long long ld = d.hpl_ld;
        const int p = d.lm_order_pose[k0 + i];
        red_add(g + 3 * p + 0, B[0] * u0 + B[ld] * u1);
        red_add(g + 3 * p + 1, B[2 * ld] * u0 + B[3 * ld] * u1);
        red_add(g + 3 * p + 2, B[4 * ld] * u0 + B[5 * ld] * u1);
    }
    const long long npairs = (long long)m * (m + 1) / 2;
    for (long long idx = lane; idx < npairs; idx += 32) {
        int i = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
        while ((long long)i * (i + 1) / 2 > idx) i--;
        while ((long long)(i + 1) * (i + 2) / 2 <= idx) i++;
        const int j = (int)(idx - (long long)i * (i + 1) / 2);
        const S* Bi = d.Hpl + d.lm_order[k0 + i];
        const S* Bj = d.Hpl + d.lm_order[k0 + j];
        const long long ld = d.hpl_ld;
        const int pi = d.lm_order_pose[k0 + i], pj = d.lm_order_pose[k0 + j];  // ascending pose inside a landmark: pi >= pj
        S bi[6], bj[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { bi[k] = Bi[k * ld]; bj[k] = Bj[k * ld]; }
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const S ya0 = bi[2 * a] * i00 + bi[2 * a + 1] * i01;
            const S ya1 = bi[2 * a] * i01 + bi[2 * a + 1] * i11;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (i == j && c > a) continue;  // diagonal block: lower part only
                red_add(Sm + (size_t)(3 * pi + a) + (size_t)(3 * pj + c) * ldS, -(ya0 * bj[2 * c] + ya1 * bj[2 * c + 1]));
            }
        }
    }
}

// ---- blocked Cholesky ---------------------------------------------------------------------------------
// diagonal block: unblocked right-looking Cholesky in shared memory, one CTA, ONE barrier per column.  Thread (r, q) = (tid / 4,
// tid % 4) updates the entries c = j + 1 + q, j + 5 + q, ... of row r with the UNSCALED column j:
// A[r][c] -= A[r][j] A[c][j] / A[j][j]; column j itself is scaled by 1 / sqrt(A[j][j]) after the barrier (nobody reads it again).
template <typename S>
__global__ void __launch_bounds__(256) k_potrf_diag(S* __restrict__ Sm, int ld, int k0, int kb, double* __restrict__ stats) {
    __shared__ S A[NB][NB + 1];  // A[row][col]
    for (int t = threadIdx.x; t < kb * kb; t += blockDim.x) {
        int c = t / kb, r = t % kb;
        A[r][c] = (r >= c) ? Sm[(size_t)(k0 + r) + (size_t)(k0 + c) * ld] : S(0);
    }
    __syncthreads();
    const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
    S sprev = S(1);
    for (int j = 0; j < kb; j++) {
        S p = A[j][j];
        if (!(p > S(0))) { if (threadIdx.x == 0) stats[5] = 1.0; p = (p < S(0)) ? -p : S(1e-30); }
        const S s = (S)rsqrt((double)p);
        if (j > 0 && q == 0 && r >= j - 1 && r < kb) A[r][j - 1] *= sprev;     // finish column j - 1 (diagonal: p_prev * s_prev = sqrt)
        if (r > j && r < kb) {
            const S f = A[r][j] * (s * s);
            int c = j + 1 + q;
            for (; c + 12 <= r; c += 16) {   // four independent updates in flight
                const S a0 = A[c][j], a1 = A[c + 4][j], a2 = A[c + 8][j], a3 = A[c + 12][j];
                const S b0 = A[r][c], b1 = A[r][c + 4], b2 = A[r][c + 8], b3 = A[r][c + 12];
                A[r][c] = b0 - f * a0; A[r][c + 4] = b1 - f * a1; A[r][c + 8] = b2 - f * a2; A[r][c + 12] = b3 - f * a3;
            }
            for (; c <= r; c += 4) A[r][c] -= f * A[c][j];
        }
        sprev = s;
        __syncthreads();
    }
    if (q == 0 && r == kb - 1) A[r][kb - 1] *= sprev;
    __syncthreads();
    for (int t = threadIdx.x; t < kb * kb; t += blockDim.x) {
        int c = t / kb, rr = t % kb;
        if (rr >= c) Sm[(size_t)(k0 + rr) + (size_t)(k0 + c) * ld] = A[rr][c];
    }
}

// panel: A[i, k0:k0+kb] <- A[i, k0:k0+kb] * L_kk^-T, one thread per row below the diagonal block.  Column-oriented: once x[m] is
// final, the updates of x[m+1..] are independent FMAs (the dependent chain is one multiply per column, not the whole dot product).
template <typename S>
__global__ void __launch_bounds__(128) k_trsm_panel(S* __restrict__ Sm, int ld, int nend, int k0, int kb) {
    __shared__ S L[NB][NB + 1];  // L[row j][col m]
    __shared__ S rd[NB];
    for (int t = threadIdx.x; t < NB * NB; t += blockDim.x) {
        int j = t % NB, m = t / NB;   // consecutive threads walk down a column: coalesced
        S v = (j == m) ? S(1) : S(0);
        if (j < kb && m <= j) v = Sm[(size_t)(k0 + j) + (size_t)(k0 + m) * ld];
        L[j][m] = v;
    }
    __syncthreads();
    if (threadIdx.x < NB) rd[threadIdx.x] = S(1) / L[threadIdx.x][threadIdx.x];
    __syncthreads();
    const int i = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nend) return;
    S x[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) x[j] = (j < kb) ? Sm[(size_t)i + (size_t)(k0 + j) * ld] : S(0);
#pragma unroll
    for (int m = 0; m < NB; m++) {
        x[m] *= rd[m];
#pragma unroll
        for (int j = m + 1; j < NB; j++) x[j] -= x[m] * L[j][m];
    }
#pragma unroll
    for (int j = 0; j < NB; j++)
        if (j < kb) Sm[(size_t)i + (size_t)(k0 + j) * ld] = x[j];
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// trailing update: C[ti,tj] -= P[ti] P[tj]^T, 64x64 per CTA, P = the kt columns k0..k0+kt of the factor (kt a multiple of 64
// except for the last panel), accumulated over 64-wide chunks of k.  Updated region: rows r0.., columns r0..r0+64*TC (TC =
// number of column tiles; TC >= T means the whole lower triangle beyond r0), lower tiles only (ti >= tj).
// 128 threads = 4 warps in a 2x2 arrangement of 32x32 warp tiles; FP64 on the tensor pipe (DMMA m8n8k4).
// Two-level blocking: inside an outer panel of kOuter columns only the panel's own remaining columns are updated after each
// 64-wide step (narrow region), the rest of the matrix once per outer panel with kt = kOuter: the trailing matrix is read
// and written kOuter/64 times less often, which turns the update from HBM bound into tensor-pipe bound.
template <typename S>
__global__ void __launch_bounds__(128) k_syrk_tiles(S* __restrict__ Sm, int ld, int nend, int k0, int kt, int r0, int T, int TC) {
    extern __shared__ unsigned char smem_raw[];
    S* sA = reinterpret_cast<S*>(smem_raw);   // [NB k][LDT rows]
    S* sB = sA + NB * LDT;
    int ti, tj;
    if (TC >= T) {   // triangular decode of a 1-D grid
        const long long idx = blockIdx.x;
        ti = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
        while ((long long)ti * (ti + 1) / 2 > idx) ti--;
        while ((long long)(ti + 1) * (ti + 2) / 2 <= idx) ti++;
        tj = (int)(idx - (long long)ti * (ti + 1) / 2);
    } else {         // narrow region: grid (T, TC)
        ti = blockIdx.x; tj = blockIdx.y;
        if (ti < tj) return;
    }
    const int ra = r0 + ti * NB, rb = r0 + tj * NB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    double accd[4][4][2];
    float accf[8][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) accd[a][c][0] = accd[a][c][1] = 0.0;
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) accf[a][c] = 0.f;
    for (int kc = 0; kc < kt; kc += NB) {
        const int kb = (kt - kc < NB) ? kt - kc : NB;
        if (kc) __syncthreads();
        for (int t = threadIdx.x; t < NB * NB; t += blockDim.x) {
            const int k = t / NB, r = t % NB;
            S va = S(0), vb = S(0);
            if (k < kb) {
                if (ra + r < nend) va = Sm[(size_t)(ra + r) + (size_t)(k0 + kc + k) * ld];
                if (rb + r < nend) vb = Sm[(size_t)(rb + r) + (size_t)(k0 + kc + k) * ld];
            }
            sA[k * LDT + r] = va;
            sB[k * LDT + r] = vb;
        }
        __syncthreads();
        if constexpr (sizeof(S) == 8) {
#pragma unroll 4
            for (int kk = 0; kk < NB; kk += 4) {
                double af[4], bf[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    af[t] = sA[(kk + fk) * LDT + wm + t * 8 + fr];
                    bf[t] = sB[(kk + fk) * LDT + wn + t * 8 + fr];
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) dmma_m8n8k4(accd[a][c][0], accd[a][c][1], af[a], bf[c]);
            }
        } else {
            for (int k = 0; k < NB; k++) {
                float av[8], bv[4];
#pragma unroll
                for (int a = 0; a < 8; a++) av[a] = sA[k * LDT + tx * 8 + a];
#pragma unroll
                for (int c = 0; c < 4; c++) bv[c] = sB[k * LDT + ty * 4 + c];
#pragma unroll
                for (int a = 0; a < 8; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) accf[a][c] += av[a] * bv[c];
            }
        }
    }
    if constexpr (sizeof(S) == 8) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int row = ra + wm + a * 8 + fr;
                const int col = rb + wn + c * 8 + 2 * fk;
                if (row < nend) {   // lower triangle only: in skyline storage an entry above the diagonal is somebody else's
                    if (col <= row) Sm[(size_t)row + (size_t)col * ld] -= accd[a][c][0];
                    if (col + 1 <= row) Sm[(size_t)row + (size_t)(col + 1) * ld] -= accd[a][c][1];
                }
            }
    } else {
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int row = ra + tx * 8 + a, col = rb + ty * 4 + c;
                if (row < nend && col <= row) Sm[(size_t)row + (size_t)col * ld] -= accf[a][c];
            }
    }
}

// The bulk trailing update (once per outer panel, kt = kOuter): 128 x 64 tiles, 256 threads = 8 warps (4 x 2) of 32 x 32
// DMMA warp tiles, k in chunks of 32 through a 2-stage cp.async pipeline (the next chunk streams into shared memory while the
// tensor pipe works on the current one).  Rows [r0, n) x columns [r0, n), lower part only: tile (ti, tj) is needed when
// 128 ti + 127 >= 64 tj; entries above the diagonal of a straddling tile are computed but not stored.
constexpr int KC = 32;           // k chunk
constexpr int LDA2 = 128 + 4;    // == 4 (mod 16) doubles: conflict-free fragment loads
constexpr int LDB2 = 64 + 4;
__device__ __forceinline__ void cp_async8(void* dst, const void* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 8 : 0;   // src-size 0: zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__global__ void __launch_bounds__(256, 2) k_syrk_big(double* __restrict__ Sm, int ld, int nend, int k0, int kt, int r0) {
    const int ti = blockIdx.x, tj = blockIdx.y;
    if (tj > 2 * ti + 1) return;
    extern __shared__ unsigned char smem_raw[];
    double* sbuf = reinterpret_cast<double*>(smem_raw);
    constexpr int kStage = KC * LDA2 + KC * LDB2;
    const int ra = r0 + ti * 128, rb = r0 + tj * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    auto load_stage = [&](int stg, int kc) {
        double* sA = sbuf + stg * kStage;
        double* sB = sA + KC * LDA2;
#pragma unroll
        for (int j = 0; j < KC * 128 / 256; j++) {
            const int idx = tid + 256 * j, r = idx & 127, k = idx >> 7;
            const bool ok = (ra + r < nend) && (kc + k < kt);
            cp_async8(sA + k * LDA2 + r, Sm + (ok ? (size_t)(ra + r) + (size_t)(k0 + kc + k) * ld : 0), ok);
        }
#pragma unroll
        for (int j = 0; j < KC * 64 / 256; j++) {
            const int idx = tid + 256 * j, r = idx & 63, k = idx >> 6;
            const bool ok = (rb + r < nend) && (kc + k < kt);
            cp_async8(sB + k * LDB2 + r, Sm + (ok ? (size_t)(rb + r) + (size_t)(k0 + kc + k) * ld : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c][0] = acc[a][c][1] = 0.0;
    const int nch = (kt + KC - 1) / KC;
    load_stage(0, 0);
    for (int ch = 0; ch < nch; ch++) {
        if (ch + 1 < nch) {
            load_stage((ch + 1) & 1, (ch + 1) * KC);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* sA = sbuf + (ch & 1) * kStage;
        const double* sB = sA + KC * LDA2;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                af[t] = sA[(kk + fk) * LDA2 + wm + t * 8 + fr];
                bf[t] = sB[(kk + fk) * LDB2 + wn + t * 8 + fr];
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) dmma_m8n8k4(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
        }
        __syncthreads();   // the stage is refilled two iterations later
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int row = ra + wm + a * 8 + fr;
            const int col = rb + wn + c * 8 + 2 * fk;
            if (row < nend) {
                if (row >= col) Sm[(size_t)row + (size_t)col * ld] -= acc[a][c][0];
                if (row >= col + 1) Sm[(size_t)row + (size_t)(col + 1) * ld] -= acc[a][c][1];
            }
        }
}

// ---- triangular solves with the factor ------------------------------------------------------------------
// forward, diagonal block: y_k = L_kk^-1 g_k   (one CTA of NB threads)
template <typename S>
__global__ void __launch_bounds__(NB) k_fwd_diag(const S* __restrict__ Sm, S* __restrict__ g, int ld, int k0, int kb) {
    __shared__ S y[NB];
    const int j = threadIdx.x;
    S v = (j < kb) ? g[k0 + j] : S(0);
    for (int m = 0; m < kb; m++) {
        if (j == m) y[m] = v / Sm[(size_t)(k0 + m) + (size_t)(k0 + m) * ld];
        __syncthreads();
        if (j > m && j < kb) v -= Sm[(size_t)(k0 + j) + (size_t)(k0 + m) * ld] * y[m];
    }
    if (j < kb) g[k0 + j] = y[j];
}
// forward, below the block: g_i -= sum_m L[i, k0+m] y_m
template <typename S>
__global__ void __launch_bounds__(256) k_fwd_update(const S* __restrict__ Sm, S* __restrict__ g, int ld, int nend, int k0, int kb) {
    __shared__ S y[NB];
    if (threadIdx.x < NB) y[threadIdx.x] = (threadIdx.x < kb) ? g[k0 + threadIdx.x] : S(0);
    __syncthreads();
    const int i = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nend) return;
    S s = S(0);
    for (int m = 0; m < kb; m++) s += Sm[(size_t)i + (size_t)(k0 + m) * ld] * y[m];
    g[i] -= s;
}
// backward, gather: t_m = sum_{i >= k0+kb} L[i, k0+m] x_i   (t zeroed by the caller)
template <typename S>
__global__ void __launch_bounds__(256) k_bwd_gather(const S* __restrict__ Sm, const S* __restrict__ x, S* __restrict__ t,
                                                    int ld, int nend, int k0, int kb) {
    __shared__ S part[8][NB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
    const S xi = (i < nend) ? x[i] : S(0);
    for (int m = 0; m < kb; m++) {
        S v = (i < nend) ? Sm[(size_t)i + (size_t)(k0 + m) * ld] * xi : S(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BOS_FULL_MASK, v, o);
        if (lane == 0) part[warp][m] = v;
    }
    __syncthreads();
    if (threadIdx.x < kb) {
        S s = S(0);
        for (int w = 0; w < 8; w++) s += part[w][threadIdx.x];
        red_add(t + threadIdx.x, s);
    }
}
// backward, diagonal block: x_k = L_kk^-T (y_k - t)
template <typename S>
__global__ void __launch_bounds__(NB) k_bwd_diag(const S* __restrict__ Sm, S* __restrict__ g, const S* __restrict__ t, int ld, int k0, int kb) {
    __shared__ S x[NB];
    const int j = threadIdx.x;
    S v = (j < kb) ? g[k0 + j] - t[j] : S(0);
    for (int m = kb - 1; m >= 0; m--) {
        if (j == m) x[m] = v / Sm[(size_t)(k0 + m) + (size_t)(k0 + m) * ld];
        __syncthreads();
        if (j < m) v -= Sm[(size_t)(k0 + m) + (size_t)(k0 + j) * ld] * x[m];  // L^T[j][m] = L[m][j]
    }
    if (j < kb) g[k0 + j] = x[j];
}

template <typename S>
__global__ void __launch_bounds__(256) k_copy_delta_p(const S* __restrict__ g, S* __restrict__ delta, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) delta[i] = g[i];
}

// Blocked Cholesky of the lower triangle of a column-major matrix, in place (the strict upper triangle is never touched).
// Dense: ld = n, panel_end = nullptr.  Skyline: column j holds rows j .. < panel_end[j / NB] (a non-decreasing row limit per 64-column panel,
// from the cached symbolic phase) in storage with leading dimension ld = W - 1, W rows per column (element (i, j) at A[i + j * ld], the LAPACK
// band-storage trick): every kernel below sees an ordinary column-major window, bounded by the panel's row limit.
// stats[5] is set when a pivot is not positive.  Returns the number of launches.
static bool dbg_sync_on() { static int v = -1; if (v < 0) v = std::getenv("BOS_DEBUG_SYNC") ? 1 : 0; return v == 1; }
#define DBG_SYNC(what, a, b) do { if (dbg_sync_on()) { cudaError_t e__ = cudaStreamSynchronize(st); if (e__ != cudaSuccess) { std::fprintf(stderr, "[dense] %s (%d, %d): %s\n", what, (int)(a), (int)(b), cudaGetErrorString(e__)); return -1; } } } while (0)
template <typename S>
int skyline_cholesky_lower(S* Smat, int n, int ld, const int* panel_end, double* stats, cudaStream_t st) {
    int nl = 0;
    const size_t smem = 2 * (size_t)NB * LDT * sizeof(S);
    ensure_dyn_smem((const void*)k_syrk_tiles<S>, smem);
    auto rend_of = [&](int k0) { return panel_end ? panel_end[k0 / NB] : n; };
    for (int c0 = 0; c0 < n; c0 += kDenseOuter) {
        const int cend = (c0 + kDenseOuter < n) ? c0 + kDenseOuter : n;
        for (int k0 = c0; k0 < cend; k0 += NB) {
            const int kb = (cend - k0 < NB) ? cend - k0 : NB;
            k_potrf_diag<S><<<1, 256, 0, st>>>(Smat, ld, k0, kb, stats); nl++;
            DBG_SYNC("potrf", k0, kb);
            const int r0 = k0 + kb, rend = rend_of(k0);
            if (r0 < rend) {
                k_trsm_panel<S><<<(rend - r0 + 127) / 128, 128, 0, st>>>(Smat, ld, rend, k0, kb); nl++;
                DBG_SYNC("trsm", k0, rend);
                if (r0 < cend) {   // the outer panel's own remaining columns
                    const int T = (rend - r0 + NB - 1) / NB, TC = (cend - r0 + NB - 1) / NB;
                    if (TC >= T) k_syrk_tiles<S><<<(unsigned)((long long)T * (T + 1) / 2), 128, smem, st>>>(Smat, ld, rend, k0, kb, r0, T, T);
                    else k_syrk_tiles<S><<<dim3(T, TC), 128, smem, st>>>(Smat, ld, rend, k0, kb, r0, T, TC);
                    nl++;
                    DBG_SYNC("syrk narrow", k0, rend);
                }
            }
        }
        const int wend = rend_of(cend - 1);   // rows the outer panel's columns reach (non-decreasing limits: the last panel's)
        if (cend < wend) {            // everything beyond the outer panel, once, with all of its columns
            const int T = (wend - cend + NB - 1) / NB;
            if constexpr (sizeof(S) == 8) {
                constexpr size_t smem_big = 2 * (size_t)(KC * LDA2 + KC * LDB2) * sizeof(double);
                ensure_dyn_smem((const void*)k_syrk_big, smem_big);
                k_syrk_big<<<dim3((wend - cend + 127) / 128, T), 256, smem_big, st>>>(Smat, ld, wend, c0, cend - c0, cend);
            } else {
                k_syrk_tiles<S><<<(unsigned)((long long)T * (T + 1) / 2), 128, smem, st>>>(Smat, ld, wend, c0, cend - c0, cend, T, T);
            }
            nl++;
            DBG_SYNC("syrk bulk", c0, wend);
        }
    }
    return nl;
}
template <typename S>
int dense_cholesky_lower(S* Smat, int n, double* stats, cudaStream_t st) { return skyline_cholesky_lower<S>(Smat, n, n, nullptr, stats, st); }
template int dense_cholesky_lower<double>(double*, int, double*, cudaStream_t);
template int dense_cholesky_lower<float>(float*, int, double*, cudaStream_t);

// Schur complement of the landmark blocks into a dense (ld = n) or skyline (w.sky: ld = W - 1, W rows per column) lower triangle, Cholesky,
// the two triangular solves, landmark back-substitution
template <typename S>
static int dense_solve_eager(const Dev<S>& d, DenseWork<S>& w, cudaStream_t st, int* launches);

template <typename S>
int launch_dense_solve(const Dev<S>& d, DenseWork<S>& w, double damping, cudaStream_t st, int* launches) {
    (void)damping;
    static const bool no_graph = std::getenv("BOS_NO_GRAPH") != nullptr;
    // what the captured kernels were given: the device view and the work buffers
    std::vector<unsigned char> key(sizeof(Dev<S>) + 6 * sizeof(void*) + 3 * sizeof(int));
    {
        unsigned char* k = key.data();
        std::memcpy(k, &d, sizeof(Dev<S>)); k += sizeof(Dev<S>);
        const void* ptrs[6] = {w.Smat, w.g, w.hllinv, w.ul, w.tl, w.tl_blk};
        std::memcpy(k, ptrs, sizeof(ptrs)); k += sizeof(ptrs);
        const int ints[3] = {w.n, w.sky ? 1 : 0, w.sky_W};
        std::memcpy(k, ints, sizeof(ints));
    }
    if (w.graph && key != w.graph_key) { cudaGraphExecDestroy(w.graph); w.graph = nullptr; w.eager_calls = 0; w.graph_failed = false; }
    if (w.graph) {
        if (cudaGraphLaunch(w.graph, st) != cudaSuccess) return -1;
        if (launches) *launches = w.graph_launches;
        return 0;
    }
    if (!no_graph && !dbg_sync_on() && !w.graph_failed && w.eager_calls >= 1) {
        // the first solve ran eagerly (function attributes set, modules loaded); capture this one
        int nl = 0;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int rc = dense_solve_eager<S>(d, w, st, &nl);
            cudaGraph_t g = nullptr;
            const cudaError_t e = cudaStreamEndCapture(st, &g);
            cudaGraphExec_t ge = nullptr;
            if (rc == 0 && e == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) {
                cudaGraphDestroy(g);
                w.graph = ge; w.graph_launches = nl; w.graph_key = key;
                if (cudaGraphLaunch(w.graph, st) != cudaSuccess) return -1;
                if (launches) *launches = nl;
                return 0;
            }
            if (g) cudaGraphDestroy(g);
        }
        cudaGetLastError();
        w.graph_failed = true;      // stay eager
    }
    w.eager_calls++;
    return dense_solve_eager<S>(d, w, st, launches);
}

template <typename S>
static int dense_solve_eager(const Dev<S>& d, DenseWork<S>& w, cudaStream_t st, int* launches) {
    int nl = 0;
    const int n = w.n;
    const bool sky = w.sky;
    const int ld = sky ? w.sky_W - 1 : n;
    const int* pend = sky ? w.sky_panel_end.data() : nullptr;
    auto rend_of = [&](int k0) { return pend ? pend[k0 / NB] : n; };
    const int gp = (d.NP + 255) / 256, gl = (d.NL + 255) / 256;
    cudaMemsetAsync(w.Smat, 0, sizeof(S) * (sky ? (size_t)n * w.sky_W : (size_t)n * n), st);
    k_dense_fill_diag<S><<<gp, 256, 0, st>>>(d, w.Smat, w.g, ld); nl++;
    if (d.n_off > 0) { k_dense_fill_off<S><<<(d.n_off + 255) / 256, 256, 0, st>>>(d, w.Smat, ld); nl++; }
    if (d.NL > 0) {
        k_lm_prep<S><<<gl, 256, 0, st>>>(d, w.hllinv, w.ul); nl++;
        k_dense_schur<S><<<(d.NL * 32 + 255) / 256, 256, 0, st>>>(d, w.hllinv, w.ul, w.Smat, w.g, ld); nl++;
    }
    DBG_SYNC("schur fill", n, ld);
    {
        const int nc = skyline_cholesky_lower<S>(w.Smat, n, ld, pend, d.stats, st);
        if (nc < 0) return -1;
        nl += nc;
    }
    // forward substitution L y = g
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int kb = (n - k0 < NB) ? n - k0 : NB;
        k_fwd_diag<S><<<1, NB, 0, st>>>(w.Smat, w.g, ld, k0, kb); nl++;
        const int r0 = k0 + kb, rend = rend_of(k0);
        if (r0 < rend) { k_fwd_update<S><<<(rend - r0 + 255) / 256, 256, 0, st>>>(w.Smat, w.g, ld, rend, k0, kb); nl++; }
    }
    // backward substitution L^T x = y
    const int last = ((n - 1) / NB) * NB;
    for (int k0 = last; k0 >= 0; k0 -= NB) {
        const int kb = (n - k0 < NB) ? n - k0 : NB;
        const int r0 = k0 + kb, rend = rend_of(k0);
        cudaMemsetAsync(w.tl_blk, 0, sizeof(S) * NB, st);
        if (r0 < rend) { k_bwd_gather<S><<<(rend - r0 + 255) / 256, 256, 0, st>>>(w.Smat, w.g, w.tl_blk, ld, rend, k0, kb); nl++; }
        k_bwd_diag<S><<<1, NB, 0, st>>>(w.Smat, w.g, w.tl_blk, ld, k0, kb); nl++;
    }
    DBG_SYNC("triangular solves", n, ld);
    k_copy_delta_p<S><<<(n + 255) / 256, 256, 0, st>>>(w.g, d.delta, n); nl++;
    if (d.NL > 0) {
        cudaMemsetAsync(w.tl, 0, sizeof(S) * 2 * (size_t)d.NL, st);
        if (d.n_hpl > 0) {
            k_lm_gather<S><<<(d.n_hpl + 255) / 256, 256, 0, st>>>(d.n_hpl, d.Hpl, d.hpl_ld, d.lm_order, d.lm_order_pose, d.lm_order_lm,
                                                                    d.delta, w.tl, nullptr);
            nl++;
        }
        k_lm_backsub<S><<<gl, 256, 0, st>>>(d, w.hllinv, w.tl); nl++;
    }
    if (launches) *launches = nl;
    return 0;
}

template int launch_dense_solve<double>(const Dev<double>&, DenseWork<double>&, double, cudaStream_t, int*);
template int launch_dense_solve<float>(const Dev<float>&, DenseWork<float>&, double, cudaStream_t, int*);

}  // namespace bos
