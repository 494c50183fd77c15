// setup.cu -- on-device problem setup: what the reference does with std::map lookups per edge per iteration
// (framework/state.cpp:20-67) and what pattern.cpp does once on the host, as GPU sort / scan / run-length passes:
//   * id -> stix resolution of every edge end point (framework/state.cpp:43-63: map::at semantics, unknown ids are an error,
//     a duplicated pose id resolves to its LAST insertion; landmarks: ascending id, slam/triangulation.cpp:68-73)
//   * the (pose, landmark)-sorted bearing edge order with ties in caller order, the unique pose-landmark block slots, the
//     CSR-of-blocks row pointers of poses and landmarks, the landmark-major slot order and the triangulation rows
// Every table is bit-identical to the host builder's (tests compare the checksum over all pattern tables).
// The sorts and scans are CUB device primitives (setup path, not the per-iteration hot path); the packing, run-length, scatter and
// binary-search kernels are ours.
#include "bos_internal.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdio>

namespace bos {

namespace {

struct DevBuf {   // scratch of one setup call: ONE device allocation, carved up (a cudaMalloc per table costs more than the sorts themselves)
    unsigned char* base = nullptr;
    size_t cap = 0, used = 0;
    std::vector<void*> extra;
    explicit DevBuf(size_t bytes) {
        if (cudaMalloc(reinterpret_cast<void**>(&base), bytes) == cudaSuccess) cap = bytes;
        else { base = nullptr; cudaGetLastError(); }
    }
    template <typename T>
    T* get(size_t n) {
        const size_t bytes = (sizeof(T) * (n > 0 ? n : 1) + 255) / 256 * 256;
        if (base && used + bytes <= cap) { T* p = reinterpret_cast<T*>(base + used); used += bytes; return p; }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
        extra.push_back(p);
        return static_cast<T*>(p);
    }
    ~DevBuf() { if (base) cudaFree(base); for (void* p : extra) cudaFree(p); }
};

int bits_for(int n) { int b = 1; while ((1LL << b) < (long long)n) b++; return b; }

__global__ void k_pack_keys(int E, int shift, const int* __restrict__ pose, const int* __restrict__ lm, unsigned long long* __restrict__ key, int* __restrict__ idx,
                            int NP, int NL, int* __restrict__ bad) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int p = pose[e], l = lm[e];
    if (p < 0 || p >= NP || l < 0 || l >= NL) { *bad = 1; key[e] = 0; idx[e] = e; return; }
    key[e] = ((unsigned long long)(unsigned)p << shift) | (unsigned)l;
    idx[e] = e;
}
// sorted keys -> sorted (pose, lm), run heads, inverse permutation
__global__ void k_unpack_sorted(int E, int shift, const unsigned long long* __restrict__ key, const int* __restrict__ perm, int* __restrict__ pose, int* __restrict__ lm,
                                int* __restrict__ head, int* __restrict__ inv) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= E) return;
    const unsigned long long v = key[k];
    pose[k] = (int)(v >> shift);
    lm[k] = (int)(v & ((1ull << shift) - 1ull));
    head[k] = (k == 0 || key[k - 1] != v) ? 1 : 0;
    inv[perm[k]] = k;
}
// b_slot = (inclusive scan of heads) - 1; the heads scatter their (pose, lm) to the slot arrays
__global__ void k_slots(int E, const int* __restrict__ head, const int* __restrict__ scan_incl, const int* __restrict__ pose, const int* __restrict__ lm, int* __restrict__ b_slot,
                        int* __restrict__ slot_pose, int* __restrict__ slot_lm) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= E) return;
    const int s = scan_incl[k] - 1;
    b_slot[k] = s;
    if (head[k]) { slot_pose[s] = pose[k]; slot_lm[s] = lm[k]; }
}
__global__ void k_histogram(int n, const int* __restrict__ key, int* __restrict__ cnt) {   // cnt[key + 1]++
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(cnt + key[i] + 1, 1);
}
__global__ void k_iota(int n, int* __restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
__global__ void k_gather(int n, const int* __restrict__ idx, const int* __restrict__ src, int* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// ---- id resolution -------------------------------------------------------------------------------------------------------------------
// sorted (id, index) pairs; an id that occurs several times keeps its LAST index (std::map::operator[] overwrites, state.cpp:23,31)
__global__ void k_lookup(int n, const int* __restrict__ ids, int m, const int* __restrict__ sorted_id, const int* __restrict__ sorted_idx, int* __restrict__ out, int* __restrict__ missing) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids[i];
    int lo = 0, hi = m;                       // upper bound: first entry > id
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_id[mid] <= id) lo = mid + 1; else hi = mid; }
    if (lo == 0 || sorted_id[lo - 1] != id) { atomicExch(missing, i + 1); out[i] = -1; return; }
    out[i] = sorted_idx ? sorted_idx[lo - 1] : lo - 1;     // stable sort: the last of equal ids is the last inserted
}
__global__ void k_unique_heads(int n, const int* __restrict__ sorted, int* __restrict__ head) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || sorted[i - 1] != sorted[i]) ? 1 : 0;
}
__global__ void k_compact(int n, const int* __restrict__ sorted, const int* __restrict__ head, const int* __restrict__ scan_incl, int* __restrict__ uniq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && head[i]) uniq[scan_incl[i] - 1] = sorted[i];
}

template <typename K, typename V>
bool radix_sort_pairs(DevBuf& buf, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit, int end_bit, cudaStream_t st) {
    size_t bytes = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, begin_bit, end_bit, st) != cudaSuccess) return false;
    void* tmp = buf.get<unsigned char>(bytes);
    if (!tmp) return false;
    return cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, begin_bit, end_bit, st) == cudaSuccess;
}
bool scan_incl(DevBuf& buf, const int* in, int* out, int n, cudaStream_t st) {
    size_t bytes = 0;
    if (cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, st) != cudaSuccess) return false;
    void* tmp = buf.get<unsigned char>(bytes);
    if (!tmp) return false;
    return cub::DeviceScan::InclusiveSum(tmp, bytes, in, out, n, st) == cudaSuccess;
}
// counts at [1 .. n] -> row pointers in place (inclusive scan leaves ptr[0] = 0 untouched)
bool counts_to_ptr(DevBuf& buf, int* cnt, int n_plus_1, cudaStream_t st) { return scan_incl(buf, cnt, cnt, n_plus_1, st); }

template <typename T>
bool d2h(std::vector<T>& dst, const T* src, size_t n, cudaStream_t st) {
    dst.resize(n);
    return n == 0 || cudaMemcpyAsync(dst.data(), src, sizeof(T) * n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
}

}  // namespace

#define GRID(n) ((unsigned)(((n) + 255) / 256)), 256, 0, st

// Bearing-edge core of the pattern on the device; fills the same members of `core` that build_pattern's first phases fill.
int device_pattern_core(PatternCore& core, int NP, int NL, int64_t Eb64, const int32_t* b_pose, const int32_t* b_lm, cudaStream_t st, std::string& err) {
    core = PatternCore();
    if (NP <= 0 || NL < 0 || Eb64 < 0 || Eb64 > 0x3fffffff) { err = "bad sizes"; return 1; }
    const int Eb = (int)Eb64;
    core.valid = true;
    core.pose_ptr.assign(NP + 1, 0); core.lm_ptr.assign(NL + 1, 0); core.tri_ptr.assign(NL + 1, 0); core.epose_ptr.assign(NP + 1, 0);
    if (Eb == 0) return 0;
    DevBuf buf((size_t)Eb * 160 + ((size_t)NP + NL) * 16 + (8u << 20));
    const int shift = bits_for(NL > 1 ? NL : 2), key_bits = shift + bits_for(NP > 1 ? NP : 2);
    int* d_pose_in = buf.get<int>(Eb); int* d_lm_in = buf.get<int>(Eb);
    unsigned long long* d_key = buf.get<unsigned long long>(Eb); unsigned long long* d_key2 = buf.get<unsigned long long>(Eb);
    int* d_idx = buf.get<int>(Eb); int* d_perm = buf.get<int>(Eb); int* d_inv = buf.get<int>(Eb);
    int* d_pose = buf.get<int>(Eb); int* d_lm = buf.get<int>(Eb); int* d_head = buf.get<int>(Eb); int* d_scan = buf.get<int>(Eb);
    int* d_slot = buf.get<int>(Eb); int* d_spose = buf.get<int>(Eb); int* d_slm = buf.get<int>(Eb);
    int* d_bad = buf.get<int>(1);
    int* d_pose_ptr = buf.get<int>(NP + 1); int* d_epose_ptr = buf.get<int>(NP + 1); int* d_lm_ptr = buf.get<int>(NL + 1); int* d_tri_ptr = buf.get<int>(NL + 1);
    if (!d_pose_in || !d_lm_in || !d_key || !d_key2 || !d_idx || !d_perm || !d_inv || !d_pose || !d_lm || !d_head || !d_scan || !d_slot || !d_spose || !d_slm || !d_bad ||
        !d_pose_ptr || !d_epose_ptr || !d_lm_ptr || !d_tri_ptr) { err = "device allocation failed"; return 2; }
    bool ok = true;
    ok &= cudaMemcpyAsync(d_pose_in, b_pose, sizeof(int) * (size_t)Eb, cudaMemcpyHostToDevice, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(d_lm_in, b_lm, sizeof(int) * (size_t)Eb, cudaMemcpyHostToDevice, st) == cudaSuccess;
    cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    cudaMemsetAsync(d_pose_ptr, 0, sizeof(int) * (size_t)(NP + 1), st); cudaMemsetAsync(d_epose_ptr, 0, sizeof(int) * (size_t)(NP + 1), st);
    cudaMemsetAsync(d_lm_ptr, 0, sizeof(int) * (size_t)(NL + 1), st); cudaMemsetAsync(d_tri_ptr, 0, sizeof(int) * (size_t)(NL + 1), st);
    // (pose, lm) order, ties in caller order: LSD radix sort is stable
    k_pack_keys<<<GRID(Eb)>>>(Eb, shift, d_pose_in, d_lm_in, d_key, d_idx, NP, NL, d_bad);
    ok &= radix_sort_pairs(buf, d_key, d_key2, d_idx, d_perm, Eb, 0, key_bits, st);
    k_unpack_sorted<<<GRID(Eb)>>>(Eb, shift, d_key2, d_perm, d_pose, d_lm, d_head, d_inv);
    ok &= scan_incl(buf, d_head, d_scan, Eb, st);
    k_slots<<<GRID(Eb)>>>(Eb, d_head, d_scan, d_pose, d_lm, d_slot, d_spose, d_slm);
    int n_hpl = 0, bad = 0;
    ok &= cudaMemcpyAsync(&n_hpl, d_scan + (Eb - 1), sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { err = std::string("device pattern core: ") + cudaGetErrorString(cudaGetLastError()); return 2; }
    if (bad) { err = "bearing edge index out of range"; return 1; }
    // row pointers: poses over slots and over edges, landmarks over slots and over edges
    k_histogram<<<GRID(n_hpl)>>>(n_hpl, d_spose, d_pose_ptr);
    k_histogram<<<GRID(Eb)>>>(Eb, d_pose, d_epose_ptr);
    k_histogram<<<GRID(n_hpl)>>>(n_hpl, d_slm, d_lm_ptr);
    k_histogram<<<GRID(Eb)>>>(Eb, d_lm_in, d_tri_ptr);
    ok &= counts_to_ptr(buf, d_pose_ptr, NP + 1, st) && counts_to_ptr(buf, d_epose_ptr, NP + 1, st) && counts_to_ptr(buf, d_lm_ptr, NL + 1, st) &&
          counts_to_ptr(buf, d_tri_ptr, NL + 1, st);
    // slots grouped by landmark, ascending pose inside a landmark: stable sort of the slots by landmark
    int* d_sidx = buf.get<int>(n_hpl); int* d_lm_order = buf.get<int>(n_hpl); int* d_lmo_lm = buf.get<int>(n_hpl); int* d_lmo_pose = buf.get<int>(n_hpl);
    int* d_tri_key = buf.get<int>(Eb); int* d_tri_edge = buf.get<int>(Eb);
    if (!d_sidx || !d_lm_order || !d_lmo_lm || !d_lmo_pose || !d_tri_key || !d_tri_edge) { err = "device allocation failed"; return 2; }
    k_iota<<<GRID(n_hpl)>>>(n_hpl, d_sidx);
    ok &= radix_sort_pairs(buf, d_slm, d_lmo_lm, d_sidx, d_lm_order, n_hpl, 0, shift, st);
    k_gather<<<GRID(n_hpl)>>>(n_hpl, d_lm_order, d_spose, d_lmo_pose);
    // triangulation rows: the edges of a landmark in CALLER order, named by their sorted position (slam/triangulation.cpp:5-19)
    ok &= radix_sort_pairs(buf, d_lm_in, d_tri_key, d_inv, d_tri_edge, Eb, 0, shift, st);
    ok &= d2h(core.b_perm, d_perm, Eb, st) && d2h(core.b_pose, d_pose, Eb, st) && d2h(core.b_lm, d_lm, Eb, st) && d2h(core.b_slot, d_slot, Eb, st);
    ok &= d2h(core.slot_pose, d_spose, n_hpl, st) && d2h(core.slot_lm, d_slm, n_hpl, st);
    ok &= d2h(core.pose_ptr, d_pose_ptr, (size_t)NP + 1, st) && d2h(core.epose_ptr, d_epose_ptr, (size_t)NP + 1, st) && d2h(core.lm_ptr, d_lm_ptr, (size_t)NL + 1, st) &&
          d2h(core.tri_ptr, d_tri_ptr, (size_t)NL + 1, st);
    ok &= d2h(core.lm_order, d_lm_order, n_hpl, st) && d2h(core.lm_order_pose, d_lmo_pose, n_hpl, st) && d2h(core.lm_order_lm, d_lmo_lm, n_hpl, st);
    ok &= d2h(core.tri_edge, d_tri_edge, Eb, st);
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { err = std::string("device pattern core: ") + cudaGetErrorString(cudaGetLastError()); return 2; }
    core.slots_identity = (n_hpl == Eb);
    return 0;
}

// id -> stix for every edge end point, and the landmark id table (ascending ids of the observed landmarks)
int device_resolve_ids(int NP, const int32_t* pose_ids, int64_t Eb64, const int32_t* b_pose_id, const int32_t* b_lm_id, int64_t Eo64, const int32_t* o_src_id,
                       const int32_t* o_dst_id, int32_t* b_pose, int32_t* b_lm, int32_t* o_src, int32_t* o_dst, int32_t* lm_ids, int32_t* NL_out, cudaStream_t st,
                       std::string& err) {
    if (NP <= 0 || Eb64 < 0 || Eo64 < 0 || Eb64 > 0x3fffffff || Eo64 > 0x3fffffff) { err = "bad sizes"; return 1; }
    const int Eb = (int)Eb64, Eo = (int)Eo64;
    DevBuf buf((size_t)Eb * 64 + (size_t)Eo * 16 + (size_t)NP * 24 + (8u << 20));
    int* d_ids = buf.get<int>(NP); int* d_idx = buf.get<int>(NP); int* d_sid = buf.get<int>(NP); int* d_sidx = buf.get<int>(NP);
    int* d_missing = buf.get<int>(1);
    if (!d_ids || !d_idx || !d_sid || !d_sidx || !d_missing) { err = "device allocation failed"; return 2; }
    bool ok = cudaMemcpyAsync(d_ids, pose_ids, sizeof(int) * (size_t)NP, cudaMemcpyHostToDevice, st) == cudaSuccess;
    cudaMemsetAsync(d_missing, 0, sizeof(int), st);
    k_iota<<<GRID(NP)>>>(NP, d_idx);
    // ids are signed: flip the sign bit for the unsigned radix order through the full 32 bits of the int key type (CUB handles signed keys)
    ok &= radix_sort_pairs(buf, d_ids, d_sid, d_idx, d_sidx, NP, 0, 32, st);
    auto lookup = [&](const int32_t* ids_h, int n, int32_t* out_h, const int* table, const int* table_idx, int m) {
        if (n == 0) return true;
        int* d_in = buf.get<int>(n); int* d_out = buf.get<int>(n);
        if (!d_in || !d_out) return false;
        bool k = cudaMemcpyAsync(d_in, ids_h, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st) == cudaSuccess;
        k_lookup<<<GRID(n)>>>(n, d_in, m, table, table_idx, d_out, d_missing);
        k &= cudaMemcpyAsync(out_h, d_out, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        return k;
    };
    ok &= lookup(b_pose_id, Eb, b_pose, d_sid, d_sidx, NP);
    ok &= lookup(o_src_id, Eo, o_src, d_sid, d_sidx, NP);
    ok &= lookup(o_dst_id, Eo, o_dst, d_sid, d_sidx, NP);
    int NL = 0;
    if (Eb > 0) {   // landmark table: sorted unique ids of the bearing edges
        int* d_l = buf.get<int>(Eb); int* d_ls = buf.get<int>(Eb); int* d_v = buf.get<int>(Eb); int* d_vs = buf.get<int>(Eb);
        int* d_head = buf.get<int>(Eb); int* d_scan = buf.get<int>(Eb); int* d_uniq = buf.get<int>(Eb);
        if (!d_l || !d_ls || !d_v || !d_vs || !d_head || !d_scan || !d_uniq) { err = "device allocation failed"; return 2; }
        ok &= cudaMemcpyAsync(d_l, b_lm_id, sizeof(int) * (size_t)Eb, cudaMemcpyHostToDevice, st) == cudaSuccess;
        k_iota<<<GRID(Eb)>>>(Eb, d_v);
        ok &= radix_sort_pairs(buf, d_l, d_ls, d_v, d_vs, Eb, 0, 32, st);
        k_unique_heads<<<GRID(Eb)>>>(Eb, d_ls, d_head);
        ok &= scan_incl(buf, d_head, d_scan, Eb, st);
        k_compact<<<GRID(Eb)>>>(Eb, d_ls, d_head, d_scan, d_uniq);
        ok &= cudaMemcpyAsync(&NL, d_scan + (Eb - 1), sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok &= cudaStreamSynchronize(st) == cudaSuccess;
        if (ok) {
            int* d_out = buf.get<int>(Eb);
            if (!d_out) { err = "device allocation failed"; return 2; }
            k_lookup<<<GRID(Eb)>>>(Eb, d_l, NL, d_uniq, nullptr, d_out, d_missing);
            ok &= cudaMemcpyAsync(b_lm, d_out, sizeof(int) * (size_t)Eb, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            if (lm_ids) ok &= cudaMemcpyAsync(lm_ids, d_uniq, sizeof(int) * (size_t)NL, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        }
    }
    int missing = 0;
    ok &= cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { err = std::string("device id resolution: ") + cudaGetErrorString(cudaGetLastError()); return 2; }
    if (missing) { err = "an edge names a pose id that was never added (std::map::at would throw, framework/state.cpp:43-49)"; return 1; }
    if (NL_out) *NL_out = NL;
    return 0;
}

}  // namespace bos
