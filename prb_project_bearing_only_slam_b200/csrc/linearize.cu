// linearize.cu -- K1/K2/K3: per-edge errors + Jacobians and the block-sparse H, b assembly.
//
// Replaces the two per-edge loops and the damping of Solver::step (slam/solver.cpp:28-69) and
// error_and_jacobian x2 (slam/solver_jacobians.cpp:9-168).  The reference merges an N x N sparse
// temporary into H for every edge; here every edge adds straight into precomputed block slots:
//
//   K3  k_init_values      b = 0, diagonal blocks = damping * I (H += damping * I, solver.cpp:64-69),
//                          pose-pose blocks = 0            (one coalesced pass over the value prefix)
//   K1  k_linearize_bearing  edge-parallel over the (pose, landmark)-sorted SoA edge buffer:
//        - pose-landmark 3x2 block: owned by the edge -> staged per warp in shared memory and written
//          with coalesced stores (no atomics, no zero-init);
//        - pose 3x3 diagonal block + b_pose: warp-segmented reduction over the run of edges of one
//          pose, one RED per value per run head;
//        - landmark 2x2 diagonal block + b_lm: RED per edge (landmark runs are scattered);
//        - chi2 / over-threshold counts: warp + block reduction, one RED per CTA.
//   K2  k_linearize_odometry  J_dst = -J_src entry for entry, so one 3x3 M = J_s^T Omega J_s and one
//        3-vector serve the source block, the destination block and the off-diagonal block.
//
// The fixed pose (gauge, solver.cpp:72-73) is handled by zeroing its Jacobian blocks at the source:
// its rows/cols then hold only the damping and a zero rhs, which is the same linear system as
// deleting them (dx_fixed = 0) without any special case downstream.
#include "bos_internal.h"
#include "bos_math.cuh"

namespace bos {

template <typename S>
__global__ void __launch_bounds__(256) k_init_values(S* __restrict__ vals, int N, int NP, int NL, long long prefix_len,
                                                     S damping, double* __restrict__ stats) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (i < 8) stats[i] = 0.0;
    const long long hpp0 = N, hll0 = (long long)N + 6LL * NP, off0 = hll0 + 3LL * NL;
    for (; i < prefix_len; i += stride) {
        S v = S(0);
        if (i >= hpp0 && i < hll0) {
            int k = (int)((i - hpp0) % 6);
            if (k == 0 || k == 3 || k == 5) v = damping;
        } else if (i >= hll0 && i < off0) {
            int k = (int)((i - hll0) % 3);
            if (k == 0 || k == 2) v = damping;
        }
        vals[i] = v;
    }
}

constexpr int kLinThreads = 256;

template <typename S, bool kIdentSlots>
__global__ void __launch_bounds__(kLinThreads) k_linearize_bearing(Dev<S> d, int e_begin, int e_end, S kernel_threshold) {
    __shared__ S stage[kLinThreads / 32][6 * 32];
    __shared__ double red[2][kLinThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double chi_acc = 0.0;
    int over_acc = 0;
    const int n = e_end - e_begin;
    const int ntiles = (n + kLinThreads - 1) / kLinThreads;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int e = e_begin + tile * kLinThreads + threadIdx.x;
        const bool valid = e < e_end;
        int p = -1 - lane, l = 0;
        S J[5] = {S(0), S(0), S(0), S(0), S(0)};
        S err = S(0), om = S(0);
        if (valid) {
            p = __ldg(d.b_pose + e);
            l = __ldg(d.b_lm + e);
            S z = __ldg(d.b_z + e);
            om = __ldg(d.b_om + e);
            PoseV<S> X = load_pose<S>(d.pose, p);
            S lx, ly;
            load_lm<S>(d.lm, l, lx, ly);
            bearing_terms<S>(X, lx, ly, z, err, J);
            // threshold robust kernel: scales the ERROR only (slam/solver.cpp:37-41)
            S chi = err * om * err;
            chi_acc += (double)chi;
            if (chi > kernel_threshold) { err *= sqrt(kernel_threshold / chi); over_acc++; }
            if (p == d.fixed) { J[0] = J[1] = J[2] = S(0); }
        }
        const S w0 = J[0] * om, w1 = J[1] * om, w2 = J[2] * om;  // (J^T omega), pose part
        const S w3 = J[3] * om, w4 = J[4] * om;                  // landmark part
        // ---- pose-landmark 3x2 block -----------------------------------------------------------
        if (kIdentSlots) {
            S* sp = &stage[warp][lane * 6];
            sp[0] = w0 * J[3]; sp[1] = w0 * J[4];
            sp[2] = w1 * J[3]; sp[3] = w1 * J[4];
            sp[4] = w2 * J[3]; sp[5] = w2 * J[4];
            __syncwarp();
            const long long base = 6LL * (e_begin + tile * kLinThreads + warp * 32);
            const long long lim = 6LL * e_end;
#pragma unroll
            for (int j = 0; j < 6; j++) {
                long long gi = base + j * 32 + lane;
                if (gi < lim) d.Hpl[gi] = stage[warp][j * 32 + lane];
            }
            __syncwarp();
        } else if (valid) {
            S* hp = d.Hpl + 6LL * __ldg(d.b_slot + e);
            red_add(hp + 0, w0 * J[3]); red_add(hp + 1, w0 * J[4]);
            red_add(hp + 2, w1 * J[3]); red_add(hp + 3, w1 * J[4]);
            red_add(hp + 4, w2 * J[3]); red_add(hp + 5, w2 * J[4]);
        }
        // ---- landmark 2x2 block and b_lm ----------------------------------------------------------
        if (valid) {
            S* hl = d.Hll + 3LL * l;
            red_add(hl + 0, w3 * J[3]);
            red_add(hl + 1, w3 * J[4]);
            red_add(hl + 2, w4 * J[4]);
            S* bl = d.b + 3LL * d.NP + 2LL * l;
            red_add(bl + 0, w3 * err);
            red_add(bl + 1, w4 * err);
        }
        // ---- pose 3x3 block and b_pose: segmented reduction over runs of equal pose ---------------
        S v[9] = {w0 * J[0], w0 * J[1], w0 * J[2], w1 * J[1], w1 * J[2], w2 * J[2], w0 * err, w1 * err, w2 * err};
        const int prev = __shfl_up_sync(BOS_FULL_MASK, p, 1);
        const bool head = (lane == 0) || (prev != p);
        const unsigned heads = __ballot_sync(BOS_FULL_MASK, head);
        const unsigned after = (lane == 31) ? 0u : (heads >> (lane + 1));
        const int run_left = after ? __ffs(after) : (32 - lane);  // lanes from me to the end of my run, inclusive
        const int max_run = __reduce_max_sync(BOS_FULL_MASK, run_left);
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            if (off < max_run) {
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    S t = __shfl_down_sync(BOS_FULL_MASK, v[k], off);
                    if (off < run_left) v[k] += t;
                }
            }
        }
        if (head && valid && p != d.fixed) {
            S* hp = d.Hpp + 6LL * p;
#pragma unroll
            for (int k = 0; k < 6; k++) red_add(hp + k, v[k]);
            S* bp = d.b + 3LL * p;
            red_add(bp + 0, v[6]); red_add(bp + 1, v[7]); red_add(bp + 2, v[8]);
        }
    }
    // ---- chi2 / over-threshold ------------------------------------------------------------------------
    double c = warp_sum(chi_acc), o = warp_sum((double)over_acc);
    if (lane == 0) { red[0][warp] = c; red[1][warp] = o; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cs = 0, os = 0;
        for (int w = 0; w < kLinThreads / 32; w++) { cs += red[0][w]; os += red[1][w]; }
        if (cs != 0.0) atomicAdd(d.stats + 0, cs);
        if (os != 0.0) atomicAdd(d.stats + 2, os);
    }
}

template <typename S>
__global__ void __launch_bounds__(128) k_linearize_odometry(Dev<S> d, int e_begin, int e_end, S kernel_threshold) {
    __shared__ double red[2][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double chi_acc = 0.0;
    int over_acc = 0;
    for (int e = e_begin + blockIdx.x * blockDim.x + threadIdx.x; e < e_end; e += gridDim.x * blockDim.x) {
        const int s = __ldg(d.o_src + e), t = __ldg(d.o_dst + e);
        const PoseV<S> Xs = load_pose<S>(d.pose, s), Xd = load_pose<S>(d.pose, t);
        const size_t Eo = (size_t)d.Eo;
        S om[6];
#pragma unroll
        for (int k = 0; k < 6; k++) om[k] = __ldg(d.o_om + k * Eo + e);
        S err[3], u0, u1;
        odometry_terms<S>(Xs, Xd, __ldg(d.o_z + e), __ldg(d.o_z + Eo + e), __ldg(d.o_z + 2 * Eo + e), err, u0, u1);
        S chi = odometry_chi<S>(om, err);
        chi_acc += (double)chi;
        S scale = S(1);
        if (chi > kernel_threshold) { scale = sqrt(kernel_threshold / chi); over_acc++; }
        S M[6], v[3];
        odometry_normal_terms<S>(Xs.c, Xs.s, u0, u1, om, err, M, v, scale);
        const bool fs = (s == d.fixed), ft = (t == d.fixed);
        if (!fs) {
            S* h = d.Hpp + 6LL * s;
#pragma unroll
            for (int k = 0; k < 6; k++) red_add(h + k, M[k]);
            S* b = d.b + 3LL * s;
            red_add(b + 0, v[0]); red_add(b + 1, v[1]); red_add(b + 2, v[2]);
        }
        if (!ft) {
            S* h = d.Hpp + 6LL * t;
#pragma unroll
            for (int k = 0; k < 6; k++) red_add(h + k, M[k]);
            S* b = d.b + 3LL * t;
            red_add(b + 0, -v[0]); red_add(b + 1, -v[1]); red_add(b + 2, -v[2]);
        }
        if (!fs && !ft) {
            // H[lo][hi] += J_lo^T Omega J_hi = -M (M symmetric, so the orientation does not matter)
            S* h = d.Hoff + 9LL * __ldg(d.o_slot + e);
            red_add(h + 0, -M[0]); red_add(h + 1, -M[1]); red_add(h + 2, -M[2]);
            red_add(h + 3, -M[1]); red_add(h + 4, -M[3]); red_add(h + 5, -M[4]);
            red_add(h + 6, -M[2]); red_add(h + 7, -M[4]); red_add(h + 8, -M[5]);
        }
    }
    double c = warp_sum(chi_acc), o = warp_sum((double)over_acc);
    if (lane == 0) { red[0][warp] = c; red[1][warp] = o; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cs = 0, os = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { cs += red[0][w]; os += red[1][w]; }
        if (cs != 0.0) atomicAdd(d.stats + 1, cs);
        if (os != 0.0) atomicAdd(d.stats + 3, os);
    }
}

template <typename S>
int launch_linearize(const Dev<S>& d, const ShardRange& r, double kernel_threshold, double damping_here,
                     bool zero_hpl, int sm_count, cudaStream_t st) {
    int launches = 0;
    long long prefix = (long long)d.N + 6LL * d.NP + 3LL * d.NL + 9LL * d.n_off;
    if (zero_hpl) prefix += 6LL * d.n_hpl;
    {
        long long blocks = (prefix + 255) / 256;
        if (blocks > 148LL * 16) blocks = 148LL * 16;
        if (blocks < 1) blocks = 1;
        k_init_values<S><<<(unsigned)blocks, 256, 0, st>>>(d.vals, d.N, d.NP, d.NL, prefix, (S)damping_here, d.stats);
        launches++;
    }
    const int nb = r.b_end - r.b_begin;
    if (nb > 0) {
        int tiles = (nb + kLinThreads - 1) / kLinThreads;
        int per_sm = 0;
        if (d.b_slot == nullptr) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_linearize_bearing<S, true>, kLinThreads, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_linearize_bearing<S, false>, kLinThreads, 0);
        if (per_sm < 1) per_sm = 1;
        int grid = sm_count * per_sm;
        if (grid > tiles) grid = tiles;
        if (d.b_slot == nullptr)
            k_linearize_bearing<S, true><<<grid, kLinThreads, 0, st>>>(d, r.b_begin, r.b_end, (S)kernel_threshold);
        else
            k_linearize_bearing<S, false><<<grid, kLinThreads, 0, st>>>(d, r.b_begin, r.b_end, (S)kernel_threshold);
        launches++;
    }
    const int no = r.o_end - r.o_begin;
    if (no > 0) {
        int grid = (no + 127) / 128;
        if (grid > sm_count * 8) grid = sm_count * 8;
        k_linearize_odometry<S><<<grid, 128, 0, st>>>(d, r.o_begin, r.o_end, (S)kernel_threshold);
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

template int launch_linearize<double>(const Dev<double>&, const ShardRange&, double, double, bool, int, cudaStream_t);
template int launch_linearize<float>(const Dev<float>&, const ShardRange&, double, double, bool, int, cudaStream_t);
template int launch_edge_terms<double>(const Dev<double>&, double*, double*, double*, double*, cudaStream_t);
template int launch_edge_terms<float>(const Dev<float>&, float*, float*, float*, float*, cudaStream_t);
template int launch_update<double>(const Dev<double>&, cudaStream_t);
template int launch_update<float>(const Dev<float>&, cudaStream_t);

}  // namespace bos
