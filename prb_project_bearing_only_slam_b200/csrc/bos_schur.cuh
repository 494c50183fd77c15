// bos_schur.cuh -- pieces shared by the dense-Cholesky and the PCG solve: warp-segmented reduction,
// landmark-block inversion, landmark back-substitution.
#pragma once

#include "bos_internal.h"
#include "bos_math.cuh"

namespace bos {

// Segmented warp reduction over runs of equal adjacent keys (no ordering assumption beyond adjacency).
// On return `head` lanes hold the sum over their run.
template <typename S, int K>
__device__ __forceinline__ void warp_run_reduce(S (&v)[K], int key, int lane, bool& head) {
    const int prev = __shfl_up_sync(BOS_FULL_MASK, key, 1);
    head = (lane == 0) || (prev != key);
    const unsigned heads = __ballot_sync(BOS_FULL_MASK, head);
    const unsigned after = (lane == 31) ? 0u : (heads >> (lane + 1));
    const int run_left = after ? __ffs(after) : (32 - lane);
    const int max_run = __reduce_max_sync(BOS_FULL_MASK, run_left);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        if (off < max_run) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                S t = __shfl_down_sync(BOS_FULL_MASK, v[k], off);
                if (off < run_left) v[k] += t;
            }
        }
    }
}

// One thread per landmark: Hll^-1 (symmetric, 3 values) and optionally u = Hll^-1 * b_l.
template <typename S>
__global__ void __launch_bounds__(256) k_lm_prep(Dev<S> d, S* __restrict__ hllinv, S* __restrict__ ul) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d.NL) return;
    const S a = d.Hll[3LL * j], b = d.Hll[3LL * j + 1], c = d.Hll[3LL * j + 2];
    const S det = a * c - b * b;
    const S i00 = c / det, i01 = -b / det, i11 = a / det;
    hllinv[3LL * j] = i00; hllinv[3LL * j + 1] = i01; hllinv[3LL * j + 2] = i11;
    if (ul) {
        const S b0 = d.b[3LL * d.NP + 2LL * j], b1 = d.b[3LL * d.NP + 2LL * j + 1];
        ul[2LL * j] = i00 * b0 + i01 * b1;
        ul[2LL * j + 1] = i01 * b0 + i11 * b1;
    }
}

// Edge-parallel over the (landmark, pose)-ordered blocks: t_l += Hpl_k^T * x_pose(k), reduced per run of
// equal landmark inside the warp, one RED per run head.  blocks: either a (lm, pose)-ordered copy (order
// == nullptr) or the (pose, lm)-ordered Hpl gathered through `order`.
template <typename S>
__global__ void __launch_bounds__(256) k_lm_gather(int n_hpl, const S* __restrict__ blocks, int ld, const int* __restrict__ order,
                                                   const int* __restrict__ k_pose, const int* __restrict__ k_lm,
                                                   const S* __restrict__ x, S* __restrict__ tl, const double* done) {
    if (done && *done != 0.0) return;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < n_hpl;
    int l = -1 - lane;
    S v[2] = {S(0), S(0)};
    if (valid) {
        l = __ldg(k_lm + k);
        const int p = __ldg(k_pose + k);
        const S* B = blocks + (order ? __ldg(order + k) : k);   // SoA planes, stride ld
        const S x0 = x[3LL * p], x1 = x[3LL * p + 1], x2 = x[3LL * p + 2];
        v[0] = B[0] * x0 + B[2LL * ld] * x1 + B[4LL * ld] * x2;
        v[1] = B[(long long)ld] * x0 + B[3LL * ld] * x1 + B[5LL * ld] * x2;
    }
    bool head;
    warp_run_reduce<S, 2>(v, l, lane, head);
    if (head && valid) {
        red_add(tl + 2LL * l, v[0]);
        red_add(tl + 2LL * l + 1, v[1]);
    }
}

// dx_l = Hll^-1 * (-b_l - t_l)   with t_l = sum Hpl^T dx_p   (landmark back-substitution, K6)
template <typename S>
__global__ void __launch_bounds__(256) k_lm_backsub(Dev<S> d, const S* __restrict__ hllinv, const S* __restrict__ tl) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d.NL) return;
    const S r0 = -d.b[3LL * d.NP + 2LL * j] - tl[2LL * j];
    const S r1 = -d.b[3LL * d.NP + 2LL * j + 1] - tl[2LL * j + 1];
    const S i00 = hllinv[3LL * j], i01 = hllinv[3LL * j + 1], i11 = hllinv[3LL * j + 2];
    d.delta[3LL * d.NP + 2LL * j] = i00 * r0 + i01 * r1;
    d.delta[3LL * d.NP + 2LL * j + 1] = i01 * r0 + i11 * r1;
}

}  // namespace bos
