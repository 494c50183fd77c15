// misc.cu -- K8 per-landmark triangulation and K9 the batched whole-iteration kernel.
#include "bos_internal.h"
#include "bos_math.cuh"

#include <cfloat>

namespace bos {

template <typename S> struct Lim;
template <> struct Lim<double> { static __device__ double eps() { return DBL_EPSILON; } static __device__ double tiny() { return DBL_MIN; } };
template <> struct Lim<float> { static __device__ float eps() { return FLT_EPSILON; } static __device__ float tiny() { return FLT_MIN; } };

template <typename S>
__device__ __forceinline__ S warp_sum_s(S v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BOS_FULL_MASK, v, o);
    return v;
}

// ---- K8: triangulate_landmarks (slam/triangulation.cpp:21-62) -----------------------------------------
// One warp per landmark.  Row i of the M x 2 system: (s, -c) . l = s*px - c*py with s,c = sin/cos(theta + alpha),
// theta = t2v(pose).z, alpha = bearing.smallestAngle().  Solved like Eigen's ColPivHouseholderQR::solve:
// pivot on the larger column norm, two Householder reflectors, nonzeroPivots() decides the rank, the
// non-pivot coordinate of a rank-1 system (M = 1) stays zero.  Rows are recomputed per pass instead of
// stored, so M is unbounded.
template <typename S>
struct TriRow { S a0, a1, b; };

template <typename S>
__device__ __forceinline__ TriRow<S> tri_row(const Dev<S>& d, int e) {
    const int p = __ldg(d.b_pose + e);
    const PoseV<S> X = load_pose<S>(d.pose, p);
    const S th = smallest_angle<S>(atan2(X.s, X.c));
    const S al = smallest_angle<S>(__ldg(d.b_z + e));
    S sn, cs;
    sincos(th + al, &sn, &cs);
    return TriRow<S>{sn, -cs, sn * X.x - cs * X.y};
}

template <typename S>
__global__ void __launch_bounds__(256) k_triangulate(Dev<S> d, int* __restrict__ single_obs) {
    const int lane = threadIdx.x & 31;
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (l >= d.NL) return;
    const int e0 = d.tri_ptr[l], M = d.tri_ptr[l + 1] - e0;
    S out[2] = {S(0), S(0)};
    if (M == 0) {
        if (lane == 0) { d.lm[2LL * l] = S(0); d.lm[2LL * l + 1] = S(0); }
        return;
    }
    if (M == 1 && lane == 0 && single_obs) atomicAdd(single_obs, 1);
    // pass A: column norms
    S n0sq = S(0), n1sq = S(0);
    for (int i = lane; i < M; i += 32) {
        TriRow<S> r = tri_row<S>(d, d.tri_edge[e0 + i]);
        n0sq += r.a0 * r.a0; n1sq += r.a1 * r.a1;
    }
    n0sq = warp_sum_s<S>(n0sq); n1sq = warp_sum_s<S>(n1sq);
    const S n0 = sqrt(n0sq), n1 = sqrt(n1sq);
    const S maxn = fmax(n0, n1);
    const S th = maxn * Lim<S>::eps() / S(M);
    const S thr = th * th;
    const int piv = (n1 > n0) ? 1 : 0;  // first maximum wins ties
    const int size = (M < 2) ? 1 : 2;
    int nonzero = size;
    if (maxn * maxn < thr * S(M)) nonzero = 0;
    const S npiv = piv ? n1 : n0;
    S nq_upd = piv ? n0 : n1, nq_dir = nq_upd;
    (void)npiv;
    // pass B: first reflector on the pivot column, applied to the other column and the rhs
    TriRow<S> r0 = tri_row<S>(d, d.tri_edge[e0]);
    const S c0 = piv ? r0.a1 : r0.a0, q0 = piv ? r0.a0 : r0.a1, b0 = r0.b;
    S tail = S(0);
    for (int i = 1 + lane; i < M; i += 32) {
        TriRow<S> r = tri_row<S>(d, d.tri_edge[e0 + i]);
        const S pc = piv ? r.a1 : r.a0;
        tail += pc * pc;
    }
    tail = warp_sum_s<S>(tail);
    S beta1, tau1, den1;
    if (tail <= Lim<S>::tiny()) { tau1 = S(0); beta1 = c0; den1 = S(1); }
    else {
        beta1 = sqrt(c0 * c0 + tail);
        if (c0 >= S(0)) beta1 = -beta1;
        den1 = c0 - beta1;
        tau1 = (beta1 - c0) / beta1;
    }
    S dq = S(0), db = S(0);
    if (tau1 != S(0)) {
        for (int i = 1 + lane; i < M; i += 32) {
            TriRow<S> r = tri_row<S>(d, d.tri_edge[e0 + i]);
            const S v = (piv ? r.a1 : r.a0) / den1;
            dq += v * (piv ? r.a0 : r.a1);
            db += v * r.b;
        }
        dq = warp_sum_s<S>(dq); db = warp_sum_s<S>(db);
    }
    S tmpq = dq + q0, tmpb = db + b0;
    S q0p = q0, b0p = b0;
    if (M == 1) { q0p = q0 * (S(1) - tau1); b0p = b0 * (S(1) - tau1); }
    else if (tau1 != S(0)) { q0p = q0 - tau1 * tmpq; b0p = b0 - tau1 * tmpb; }
    S beta2 = S(1), b1pp = S(0);
    if (size == 2) {
        // transformed rows i >= 1 of the other column / rhs: q_i' = q_i - tau1 v_i tmpq, b_i' = b_i - tau1 v_i tmpb
        auto row_t = [&](int i, S& qi, S& bi) {
            TriRow<S> r = tri_row<S>(d, d.tri_edge[e0 + i]);
            const S v = (tau1 != S(0)) ? (piv ? r.a1 : r.a0) / den1 : S(0);
            qi = (piv ? r.a0 : r.a1); bi = r.b;
            if (tau1 != S(0)) { qi -= tau1 * v * tmpq; bi -= tau1 * v * tmpb; }
        };
        // norm downdate of the remaining column (decides nonzeroPivots at k = 1)
        if (nq_upd != S(0)) {
            S temp = fabs(q0p) / nq_upd;
            temp = (S(1) + temp) * (S(1) - temp);
            temp = temp < S(0) ? S(0) : temp;
            const S rr = nq_upd / nq_dir;
            const S temp2 = temp * (rr * rr);
            if (temp2 <= sqrt(Lim<S>::eps())) {
                S s2 = S(0);
                for (int i = 1 + lane; i < M; i += 32) { S qi, bi; row_t(i, qi, bi); s2 += qi * qi; }
                s2 = warp_sum_s<S>(s2);
                nq_dir = sqrt(s2); nq_upd = nq_dir;
            } else {
                nq_upd *= sqrt(temp);
            }
        }
        if (nonzero == size && nq_upd * nq_upd < thr * S(M - 1)) nonzero = 1;
        // second reflector on q'[1:]
        S c1, b1;
        row_t(1, c1, b1);
        S tail2 = S(0);
        for (int i = 2 + lane; i < M; i += 32) { S qi, bi; row_t(i, qi, bi); tail2 += qi * qi; }
        tail2 = warp_sum_s<S>(tail2);
        S tau2, den2;
        if (tail2 <= Lim<S>::tiny()) { tau2 = S(0); beta2 = c1; den2 = S(1); }
        else {
            beta2 = sqrt(c1 * c1 + tail2);
            if (c1 >= S(0)) beta2 = -beta2;
            den2 = c1 - beta2;
            tau2 = (beta2 - c1) / beta2;
        }
        b1pp = b1;
        if (nonzero == 2) {
            if (M - 1 == 1) b1pp = b1 * (S(1) - tau2);
            else if (tau2 != S(0)) {
                S db2 = S(0);
                for (int i = 2 + lane; i < M; i += 32) { S qi, bi; row_t(i, qi, bi); db2 += (qi / den2) * bi; }
                db2 = warp_sum_s<S>(db2);
                b1pp = b1 - tau2 * (db2 + b1);
            }
        }
    }
    if (nonzero == 2) {
        const S x1 = b1pp / beta2;
        const S x0 = (b0p - q0p * x1) / beta1;
        out[piv] = x0; out[1 - piv] = x1;
    } else if (nonzero == 1) {
        out[piv] = b0p / beta1;
    }
    if (lane == 0) { d.lm[2LL * l] = out[0]; d.lm[2LL * l + 1] = out[1]; }
}

template <typename S>
int launch_triangulate(const Dev<S>& d, int* single_obs_count_dev, cudaStream_t st) {
    if (d.NL > 0) k_triangulate<S><<<(d.NL * 32 + 255) / 256, 256, 0, st>>>(d, single_obs_count_dev);
    return 1;
}

// ---- K9: one GN iteration for each of nprob independent small problems in ONE launch ----------------------
// One warp per problem, one problem per CTA.  The whole iteration (slam/solver.cpp:27-97) runs out of shared
// memory: dense H (N x N), b, state; lanes stride the edges for linearization, then a warp-level Cholesky
// and two triangular solves, then boxplus.  The fixed pose is handled as in the big path (its Jacobian
// blocks are zeroed, its diagonal keeps the damping, so dx_fixed = 0).
template <typename S>
__global__ void __launch_bounds__(32) k_batch_step(BatchDev<S> d, S kernel_threshold, S damping) {
    extern __shared__ unsigned char smem_raw[];
    const int N = 3 * d.NP + 2 * d.NL;
    const int ld = N + 1;
    S* H = reinterpret_cast<S*>(smem_raw);  // [N][ld] lower triangle used, H[row][col]
    S* b = H + (size_t)N * ld;              // [N]
    S* pose = b + N;                        // [NP][4]
    S* lm = pose + 4 * d.NP;                // [NL][2]
    const int lane = threadIdx.x;
    const int prob = blockIdx.x;
    if (prob >= d.nprob) return;
    S* gpose = d.pose + (size_t)prob * d.NP * 4;
    S* glm = d.lm + (size_t)prob * d.NL * 2;
    for (int i = lane; i < N * ld; i += 32) H[i] = S(0);
    for (int i = lane; i < N; i += 32) b[i] = S(0);
    for (int i = lane; i < 4 * d.NP; i += 32) pose[i] = gpose[i];
    for (int i = lane; i < 2 * d.NL; i += 32) lm[i] = glm[i];
    __syncwarp();
    for (int i = lane; i < N; i += 32) H[i * ld + i] = damping;
    __syncwarp();
    double chi_b = 0.0, chi_o = 0.0;
    const int lmoff = 3 * d.NP;
    for (int e = lane; e < d.Eb; e += 32) {
        const int p = d.b_pose[e], l = d.b_lm[e];
        const PoseV<S> X{pose[4 * p], pose[4 * p + 1], pose[4 * p + 2], pose[4 * p + 3]};
        S err, J[5];
        bearing_terms<S>(X, lm[2 * l], lm[2 * l + 1], d.b_z[(size_t)prob * d.Eb + e], err, J);
        const S om = d.b_om[e];
        const S chi = err * om * err;
        chi_b += (double)chi;
        if (chi > kernel_threshold) err *= sqrt(kernel_threshold / chi);
        if (p == d.fixed) { J[0] = J[1] = J[2] = S(0); }
        const int idx[5] = {3 * p, 3 * p + 1, 3 * p + 2, lmoff + 2 * l, lmoff + 2 * l + 1};
#pragma unroll
        for (int a = 0; a < 5; a++) {
            const S wa = J[a] * om;
#pragma unroll
            for (int c = 0; c <= a; c++) {
                // lower triangle: row = max index.  idx is increasing in a because poses precede landmarks.
                atomicAdd(&H[idx[a] * ld + idx[c]], wa * J[c]);
            }
            atomicAdd(&b[idx[a]], wa * err);
        }
    }
    for (int e = lane; e < d.Eo; e += 32) {
        const int s = d.o_src[e], t = d.o_dst[e];
        const PoseV<S> Xs{pose[4 * s], pose[4 * s + 1], pose[4 * s + 2], pose[4 * s + 3]};
        const PoseV<S> Xd{pose[4 * t], pose[4 * t + 1], pose[4 * t + 2], pose[4 * t + 3]};
        S om[6];
#pragma unroll
        for (int k = 0; k < 6; k++) om[k] = d.o_om[6 * e + k];
        const S* z = d.o_z + ((size_t)prob * d.Eo + e) * 3;
        S err[3], u0, u1;
        odometry_terms<S>(Xs, Xd, z[0], z[1], z[2], err, u0, u1);
        const S chi = odometry_chi<S>(om, err);
        chi_o += (double)chi;
        S scale = S(1);
        if (chi > kernel_threshold) scale = sqrt(kernel_threshold / chi);
        S M[6], v[3];
        odometry_normal_terms<S>(Xs.c, Xs.s, u0, u1, om, err, M, v, scale);
        const S Mf[9] = {M[0], M[1], M[2], M[1], M[3], M[4], M[2], M[4], M[5]};
        const bool fs = (s == d.fixed), ft = (t == d.fixed);
        for (int a = 0; a < 3; a++) {
            for (int c = 0; c <= a; c++) {
                if (!fs) atomicAdd(&H[(3 * s + a) * ld + 3 * s + c], Mf[a * 3 + c]);
                if (!ft) atomicAdd(&H[(3 * t + a) * ld + 3 * t + c], Mf[a * 3 + c]);
            }
            if (!fs) atomicAdd(&b[3 * s + a], v[a]);
            if (!ft) atomicAdd(&b[3 * t + a], -v[a]);
            if (!fs && !ft) {
                const int hi = s > t ? s : t, lo = s > t ? t : s;
                for (int c = 0; c < 3; c++) atomicAdd(&H[(3 * hi + a) * ld + 3 * lo + c], -Mf[a * 3 + c]);
            }
        }
    }
    __syncwarp();
    // Cholesky H = L L^T in place (lower), columns sequential, lanes over rows
    int status = 0;
    for (int j = 0; j < N; j++) {
        S dj = H[j * ld + j];
        if (!(dj > S(0))) { status = 1; dj = (dj < S(0)) ? -dj : S(1e-30); }
        const S piv = sqrt(dj);
        __syncwarp();
        if (lane == 0) H[j * ld + j] = piv;
        for (int i = j + 1 + lane; i < N; i += 32) H[i * ld + j] /= piv;
        __syncwarp();
        for (int k = j + 1; k < N; k++) {
            const S lkj = H[k * ld + j];
            for (int i = k + lane; i < N; i += 32) H[i * ld + k] -= H[i * ld + j] * lkj;
        }
        __syncwarp();
    }
    // solve L y = -b, L^T x = y (x overwrites b)
    for (int i = lane; i < N; i += 32) b[i] = -b[i];
    __syncwarp();
    for (int j = 0; j < N; j++) {
        const S yj = b[j] / H[j * ld + j];
        __syncwarp();
        if (lane == 0) b[j] = yj;
        for (int i = j + 1 + lane; i < N; i += 32) b[i] -= H[i * ld + j] * yj;
        __syncwarp();
    }
    for (int j = N - 1; j >= 0; j--) {
        const S xj = b[j] / H[j * ld + j];
        __syncwarp();
        if (lane == 0) b[j] = xj;
        for (int i = lane; i < j; i += 32) b[i] -= H[j * ld + i] * xj;
        __syncwarp();
    }
    // boxplus
    double dinf = 0.0;
    for (int i = lane; i < d.NP; i += 32) {
        const S dx = b[3 * i], dy = b[3 * i + 1], dt = b[3 * i + 2];
        S sd, cd;
        sincos(dt, &sd, &cd);
        const S x = pose[4 * i], y = pose[4 * i + 1], c = pose[4 * i + 2], s = pose[4 * i + 3];
        gpose[4 * i] = (cd * x + (-sd) * y) + dx;
        gpose[4 * i + 1] = (sd * x + cd * y) + dy;
        gpose[4 * i + 2] = cd * c + (-sd) * s;
        gpose[4 * i + 3] = sd * c + cd * s;
        dinf = fmax(dinf, fmax(fabs((double)dx), fmax(fabs((double)dy), fabs((double)dt))));
    }
    for (int j = lane; j < d.NL; j += 32) {
        const S dx = b[lmoff + 2 * j], dy = b[lmoff + 2 * j + 1];
        glm[2 * j] = lm[2 * j] + dx;
        glm[2 * j + 1] = lm[2 * j + 1] + dy;
        dinf = fmax(dinf, fmax(fabs((double)dx), fabs((double)dy)));
    }
    chi_b = warp_sum(chi_b); chi_o = warp_sum(chi_o);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dinf = fmax(dinf, __shfl_xor_sync(BOS_FULL_MASK, dinf, o));
    if (lane == 0) {
        d.chi2[2LL * prob] = chi_b; d.chi2[2LL * prob + 1] = chi_o;
        d.delta_inf[prob] = dinf;
        d.status[prob] = status;
    }
}

size_t batch_smem_bytes(int NP, int NL, size_t scalar_bytes) {
    const size_t N = 3 * (size_t)NP + 2 * (size_t)NL;
    return (N * (N + 1) + N + 4 * (size_t)NP + 2 * (size_t)NL) * scalar_bytes;
}

template <typename S>
int launch_batch_step(const BatchDev<S>& d, double kernel_threshold, double damping, cudaStream_t st) {
    const size_t smem = batch_smem_bytes(d.NP, d.NL, sizeof(S));
    if (smem > 48 * 1024 && !ensure_dyn_smem((const void*)k_batch_step<S>, smem)) return -1;
    k_batch_step<S><<<d.nprob, 32, smem, st>>>(d, (S)kernel_threshold, (S)damping);
    return 1;
}

template int launch_triangulate<double>(const Dev<double>&, int*, cudaStream_t);
template int launch_triangulate<float>(const Dev<float>&, int*, cudaStream_t);
template int launch_batch_step<double>(const BatchDev<double>&, double, double, cudaStream_t);
template int launch_batch_step<float>(const BatchDev<float>&, double, double, cudaStream_t);

}  // namespace bos
