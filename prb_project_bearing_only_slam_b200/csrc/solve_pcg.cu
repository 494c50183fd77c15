// solve_pcg.cu -- K4/K5b/K6 for large problems: the 2x2 landmark blocks are Schur-complemented out
// IMPLICITLY and the reduced pose system  S dx_p = -(b_p - Hpl Hll^-1 b_l),
// S = Hpp - Hpl Hll^-1 Hlp, is solved by block-Jacobi preconditioned conjugate gradients.
//
// Replaces SimplicialLDLT::factorize/solve on H_nofixed (slam/solver.hpp:72, slam/solver.cpp:77-94):
// eliminating the landmark blocks is exact, so the solution is the same dx up to the PCG tolerance.
// S is never formed (at 40 observations per landmark it would be ~1600 3x3 blocks per landmark);
// one application of S is two edge-parallel passes over the pose-landmark blocks:
//   t_l  = sum_k Hpl_k^T p_pose(k)      over the (landmark, pose)-ordered copy, run-reduced per landmark
//   y_p -= sum_k Hpl_k Hll^-1 t_lm(k)   over the (pose, landmark)-ordered blocks, run-reduced per pose
// both HBM-bound streams of 6 scalars per block; vectors and landmark blocks stay L2-resident.
#include "bos_internal.h"
#include "bos_math.cuh"
#include "bos_schur.cuh"

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

template <typename S>
int launch_pcg_solve(const Dev<S>& d, PcgWork<S>& w, int max_iters, double rtol, cudaStream_t st,
                     int* iterations_out, int* launches) {
    int nl = 0;
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
