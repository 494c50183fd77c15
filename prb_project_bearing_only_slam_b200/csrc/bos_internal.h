// bos_internal.h -- shared declarations of the CUDA product library (not part of the C ABI).
#pragma once

#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <utility>

namespace bos {

// bearing edges are linearized in tiles of this many (pose, landmark)-sorted edges; the landmark-side sums of a tile
// are aggregated in shared memory through a host-precomputed tile-local grouping before they touch global memory
constexpr int kLinTile = 512;
constexpr int kDenseNB = 64;       // panel width of the blocked Cholesky
constexpr int kDenseOuter = 256;   // outer panel: the bulk of the matrix is updated once per kDenseOuter columns
// the fused PCG kernel reads the per-edge factors from two sliced-ELL layouts (see pattern.cpp)
#ifndef BOS_PCG_THREADS
#define BOS_PCG_THREADS 1024
#endif
constexpr int kPcgThreads = BOS_PCG_THREADS;       // one persistent CTA per SM
#ifndef BOS_ELL_LANES
#define BOS_ELL_LANES 8
#endif
constexpr int kEllLanesL = BOS_ELL_LANES;          // lanes per landmark row in the landmark-major layout
constexpr int kMaxPeers = 8;                        // ranks of one NVSwitch box (reduce_mode 4)
#ifndef BOS_LC_LANES
#define BOS_LC_LANES 1
#endif
constexpr int kLcLanes = BOS_LC_LANES;            // lanes per (chunk, landmark) row of the chunk-local landmark-major layout (pattern.cpp "LC")
constexpr int kShareEll = 8;                       // sharers of a (chunk, landmark) kept in the transposed list (the rest through the CSR list)

// Typed view of everything a kernel needs.  One instance per context, built after upload.
template <typename S>
struct Dev {
    int NP = 0, NL = 0, fixed = 0, Eb = 0, Eo = 0, N = 0;
    int n_hpl = 0, n_off = 0;
    int irls = 0;                 // robust kernel flavour: 0 = the reference's (scales the error only), 1 = IRLS (the weight scales Omega: H too)
    // bearing edges, sorted by (pose, landmark); SoA
    const int* b_pose = nullptr;
    const int* b_lm = nullptr;
    const S* b_z = nullptr;
    const S* b_om = nullptr;
    const int* b_slot = nullptr;  // null => slot == sorted edge index (no duplicate (pose, lm) pairs)
    const int* b_perm = nullptr;  // sorted index -> caller's edge index
    // odometry edges, caller order; o_z = [3][Eo], o_om = [6][Eo] (00 01 02 11 12 22)
    const int* o_src = nullptr;
    const int* o_dst = nullptr;
    const S* o_z = nullptr;
    const S* o_om = nullptr;
    const int* o_slot = nullptr;
    const int* oe_ptr = nullptr;     // [NP+1] odometry edges incident to a pose
    const int* oe_edge = nullptr;    // [2*Eo] (edge << 1) | role, role 0: the pose is the edge's source, 1: its destination
    const int* oe_other = nullptr;   // [2*Eo] the pose at the other end of that edge
    const int* oe2 = nullptr;        // [NP + 8][2] the first two entries of a pose's list inline: code0, code1 (-1: none)
    const unsigned char* o_shared = nullptr;  // [Eo] 1 when the edge's pose pair is shared with another edge (needs RED + zero init)
    int has_shared_off = 0;
    int Eb_pad = 0;                  // bearing SoA arrays are padded to a multiple of 4 (omega = 0 in the padding)
    int hpl_ld = 0;                  // plane stride of Hpl (multiple of 4)
    // CSR-of-blocks pattern
    const int* slot_pose = nullptr;  // [n_hpl]
    const int* slot_lm = nullptr;    // [n_hpl]
    const int* pose_ptr = nullptr;   // [NP+1] into the (pose, lm)-sorted Hpl slots
    const int* lm_ptr = nullptr;     // [NL+1] into lm_order
    const int* lm_order = nullptr;   // [n_hpl] slot indices sorted by (lm, pose)
    const int* lm_order_pose = nullptr;  // [n_hpl] pose of lm_order[k]
    const int* lm_order_lm = nullptr;    // [n_hpl] landmark of lm_order[k]
    const int* pp_ptr = nullptr;     // [NP+1] pose-pose adjacency
    const int* pp_nbr = nullptr;     // [2*n_off] neighbour pose
    const int* pp_slot = nullptr;    // [2*n_off] slot; bit 31 set when this pose is the 'hi' side (use the transpose)
    const int* off_lo = nullptr;     // [n_off] block row pose of Hoff[k]
    const int* off_hi = nullptr;     // [n_off] block column pose
    const int* epose_ptr = nullptr;  // [NP+1] CSR of the (pose, lm)-sorted bearing EDGES by pose
    const int* tile_ptr = nullptr;   // [ntiles+1] groups (distinct landmarks) of each tile
    const int* tg_lm = nullptr;      // [n_groups] landmark of the group
    const int* tg_eptr = nullptr;    // [n_groups+1] into tg_edge
    const int* tile_meta = nullptr;  // [ntiles][4] padded group offset, groups, first pose, last pose of every tile
    const int* tgp_lm = nullptr;     // tg_lm with every tile's list starting at a multiple of 8 entries (TMA alignment)
    const unsigned short* tgp_eptr = nullptr;  // tile-local tg_eptr, same offsets, groups + 1 entries per tile
    const unsigned short* tg_edge = nullptr;  // [Eb] tile-local edge index
    // sliced-ELL layouts of the bearing edges for the fused PCG kernel
    int n_clm = 0, nLg = 0;
    long long nLs = 0, nPs = 0;      // slots of the two layouts
    const int* pl_lm_id = nullptr;   // [n_clm] landmark stix of compact row r (rows sorted by descending observation count)
    const int* ell_Loff = nullptr;   // [nLg+1] column offset of each landmark group
    const int* ell_Lpose = nullptr;  // [nLs] pose of the slot; -1 for padding and for edges of the fixed pose (zero Jacobian block)
    // pose-major side: chunks of pc_cp poses (one persistent CTA each), rows sorted by edge count inside a chunk
    int pc_chunks = 0, pc_cp = 0, pc_ok = 0, pc_cl_max = 0, pc_slots_max = 0;
    const int* pc_row_pose = nullptr;      // [pc_chunks * pc_cp] pose of chunk row, -1 for padding
    const int* pc_goff = nullptr;          // [pc_chunks * pc_cp / 32 + 1] column offset of each group of 32 rows
    const int* pc_cl_ptr = nullptr;        // [pc_chunks + 1] the chunk's distinct landmark rows ...
    const int* pc_cl_row = nullptr;        // ... ascending compact landmark row ids
    const unsigned short* pc_loc = nullptr;  // [nPs] index into the chunk's landmark table, 0xffff for padding
    const int* pc_emap = nullptr;          // [nPs] sorted bearing-edge index of the slot, -1 for padding
    const int* pc_nbr = nullptr;           // [2][pc_chunks * pc_cp] first two pose-pose neighbours of the row (-1: none)
    const int* pc_nslot = nullptr;         // [2][..] their Hoff slot (bit 31: this pose is the column side)
    const int* pc_ncnt = nullptr;          // [..] number of pose-pose neighbours (more than 2: the rest through pp_ptr)
    // chunk-local landmark-major layout of the persistent PCG kernel (see pattern.cpp "LC") and the landmark sharing lists
    const int* lc_gptr = nullptr;          // [pc_chunks + 1]
    const int* lc_goff = nullptr;          // [groups + 1]
    const unsigned short* lc_row = nullptr;   // [slots] chunk-local pose row, 0xffff: none
    const unsigned short* lc_k = nullptr;     // [groups * 32] local landmark of every ELL row (rows sorted by descending edge count), 0xffff: none
    const int* sh_ptr = nullptr;           // [n_q + 1], q = pc_cl_ptr[c] + k
    const int* sh_src = nullptr;           // the q' of every chunk that sees the same landmark
    const int* sh_ell = nullptr;           // [kShareEll][n_q] the first sharers of q once more, transposed, -1 = none
    const unsigned char* sh_first = nullptr;  // [n_q] 1 in the lowest chunk that sees the landmark
    int n_q = 0;
    const int* tri_ptr = nullptr;    // [NL+1] bearing edges grouped by landmark (caller order inside a landmark)
    const int* tri_edge = nullptr;   // [Eb] sorted-edge index
    // state
    S* pose = nullptr;  // [NP][4] x,y,c,s
    S* lm = nullptr;    // [NL][2]
    S* theta = nullptr; // [NP] t2v angle of every pose, refreshed whenever the poses change (set_state, update)
    // value buffer  [ b (N) | Hpp (6 NP) | Hll (3 NL) | Hoff (9 n_off) | pad | Hpl (6 planes x hpl_ld) ]
    S* vals = nullptr;
    S* b = nullptr;
    S* Hpp = nullptr;   // xx xy xt yy yt tt
    S* Hll = nullptr;   // xx xy yy
    S* Hoff = nullptr;  // 3x3 row-major, block H[lo][hi]
    S* Hpl = nullptr;   // SoA: entry k (3x2 row-major index) of block s at Hpl[k * hpl_ld + s]
    S* Mv = nullptr;          // [9][Eo] per odometry edge: M = J_s^T Omega J_s (6) and v = J_s^T Omega e (3), written by k_linearize_odometry
    const int* cut_pose = nullptr;   // [n_cut] poses whose bearing-edge run is cut by a tile boundary (their block is assembled by REDs)
    int n_cut = 0;
    double* stats = nullptr;  // [8] chi2_b, chi2_o, over_b, over_o, delta_inf(bits), status, state digest (k_update), -
    S* delta = nullptr;       // [N]
    // fused linearize + combine over NVLink peer memory (reduce_mode 4): every rank's value buffer and statistics, own included, mapped
    // through CUDA IPC; the bearing kernel stores / REDs every block that must be combined straight into all replicas
    double* stats_k2 = nullptr;      // where the odometry kernel adds its statistics (stats, or the scratch the peer barrier publishes)
    int npeer = 0;                   // > 0: pv / pstats are valid (reduce_mode 4 and 5)
    int peer_push = 0;               // reduce_mode 4: the bearing kernel writes into every replica
    S* pv[kMaxPeers] = {};
    double* pstats[kMaxPeers] = {};
};

// tail of the value-buffer allocation (8-byte units): statistics, the odometry kernel's share of them, peer barrier slots, error flag
constexpr int kTailStats = 0, kTailStatsK2 = 8, kTailSlots = 16, kTailError = 16 + kMaxPeers, kTailWords = 64;
struct PeerBarrier {
    unsigned long long* slots[kMaxPeers];   // slots[r]: rank r's slot array (own included); rank q signals slots[r][q]
    double* pstats[kMaxPeers];
    const double* publish;                  // non-null: add these 4 statistics (chi2_o, over_o at [1], [3]) to every replica before signalling
    unsigned long long* error;              // local flag, set when the wait timed out
    int n, rank;
    unsigned long long epoch;
};
int launch_peer_barrier(const PeerBarrier& pb, cudaStream_t st);
// reduce_mode 5: after a LOCAL sharded build (mode 3's kernels) and a cross-GPU barrier, one kernel PULLS what the rank lacks from the peers'
// replicas -- the owners' pose ranges (copied in place), the landmark parts of all ranks and the few rank-boundary poses (summed in rank order
// into a scratch: the peers are still reading this rank's parts), the statistics -- and after a second barrier a commit kernel moves the scratch
// into the replica.  Bulk, coalesced NVLink reads instead of mode 4's per-scalar remote writes.
// element offset (in S) of the four double statistics inside the pull scratch: behind 5 NL landmark totals and 9 kMaxPeers boundary totals, 8-byte aligned
__host__ __device__ inline long long peer_scratch_stats_off(int NL) { return (5LL * NL + 9 * kMaxPeers + 1) & ~1LL; }
struct PeerPull {
    int n, rank, NP, NL;
    int own_p0[kMaxPeers + 1];   // pose ranges owned by the ranks' tiles
    int bnd[kMaxPeers];          // last owned pose of rank q when its bearing run continues into rank q + 1's first tile (summed from both), else -1
};
template <typename S>
int launch_peer_pull(const Dev<S>& d, const PeerPull& pp, S* scratch, int sm_count, cudaStream_t st);
template <typename S>
int launch_peer_commit(const Dev<S>& d, const PeerPull& pp, const S* scratch, cudaStream_t st);

struct ShardRange {
    int b_begin = 0, b_end = 0, o_begin = 0, o_end = 0;
};

// ---- launchers (one translation unit each) -------------------------------------------------------
template <typename S>
int launch_linearize(const Dev<S>& d, const ShardRange& r, double kernel_threshold, double damping, double damping_here,
                     bool zero_hpl, bool zero_hoff, int sm_count, cudaStream_t st, bool multi_rank, int rank, bool all_hoff, int phases = 3);
template <typename S>
int launch_edge_terms(const Dev<S>& d, S* err_b, S* jac_b, S* err_o, S* jac_o, cudaStream_t st);
template <typename S>
int launch_update(const Dev<S>& d, cudaStream_t st);
template <typename S>
int launch_pose_theta(const Dev<S>& d, cudaStream_t st);
template <typename S>
int launch_triangulate(const Dev<S>& d, int* single_obs_count_dev, cudaStream_t st);

// dense path
template <typename S>
struct DenseWork {
    S* Smat = nullptr;      // [n][n] lower triangle used, column-major (ld = n)
    S* g = nullptr;         // [n] reduced rhs / solution
    S* hllinv = nullptr;    // [NL][3]
    S* ul = nullptr;        // [NL][2]
    S* tl = nullptr;        // [NL][2]
    S* tl_blk = nullptr;    // [64] scratch of the backward substitution
    int n = 0;              // 3*NP
    size_t bytes = 0;
    // skyline flavour (BOS_SOLVER_SPARSE_CHOLESKY): Smat holds W rows per column (column j: rows j .. j + W - 1, addressed as a column-major
    // matrix with leading dimension W - 1), sky_panel_end[k] = the row limit of the 64-column panel k.  Both come from the symbolic phase
    // (skyline_symbolic), which runs once per uploaded pattern.
    bool sky = false;
    int sky_W = 0;
    std::vector<int> sky_panel_end;
    double sky_fill = 0.0;  // stored entries / (n (n + 1) / 2)
    // The launch sequence of a solve (hundreds to thousands of small kernels: one POTRF / TRSM / update per 64-column panel, one launch per
    // substitution block) depends only on the pattern: after one eager solve it is captured into a CUDA graph and replayed.  The captured
    // kernel arguments are the bytes of Dev<S> and of the work pointers: the graph is dropped when they change (new problem, robust mode).
    cudaGraphExec_t graph = nullptr;
    int graph_launches = 0, eager_calls = 0;
    bool graph_failed = false;
    std::vector<unsigned char> graph_key;
};

template <typename S>
int launch_dense_solve(const Dev<S>& d, DenseWork<S>& w, double damping, cudaStream_t st, int* launches);

// PCG path
template <typename S>
struct PcgWork {
    S* hllinv = nullptr;   // [NL][3]
    S* ul = nullptr;       // [NL][2]
    S* tl = nullptr;       // [NL][2]
    S* Hlp = nullptr;      // [n_hpl][6] copy of Hpl in (lm, pose) order
    S* minv = nullptr;     // [NP][6] inverse of the diagonal blocks of S
    S* g = nullptr;        // [3NP]
    S* x = nullptr;
    S* r = nullptr;
    S* z = nullptr;
    S* p0 = nullptr;
    S* p1 = nullptr;
    S* y = nullptr;
    // fused kernel: the operator is applied from per-EDGE factors.  A bearing edge's 3x2 block is rank one,
    // Hpl_k = Jp_k^T omega Jl_k, and Jp_k = (-j0, -j1, j0 ly - j1 lx) is determined by Jl_k = (j0, j1) and the landmark position,
    // so two scalars per edge (sqrt(omega) Jl) replace the six of the block.
    S* Lw = nullptr;               // [nLs] sqrt(omega) per landmark-major slot; only when the bearing omegas are not all equal
    S* Pw = nullptr;               // [nPs] sqrt(omega) per pose-major slot; only when the bearing omegas are not all equal
    S* Cw = nullptr;               // the same per slot of the chunk-local landmark-major layout
    S* qstat = nullptr;            // [5][n_q] per (chunk, landmark), constant during a solve: Hll^-1 (3) and the landmark position (2)
    S* tpart = nullptr;            // [2][n_q][2] per (chunk, landmark): the chunk's partial t_l, double-buffered by CG iteration parity
    int omega_uniform = 1;         // all bearing omegas equal: the pose-major pass recomputes its factors from the state
    double sqrt_omega = 1.0;       // ... with this scale
    S* hllinv_c = nullptr;         // [n_clm][3] Hll^-1 in compact row order
    S* ul4 = nullptr;              // [n_clm][4] u_l = Hll^-1 t_l (rewritten every CG iteration) and the landmark position lx, ly
    S* z4 = nullptr;               // [2][NP][4] double-buffered z (padded to one 32-byte sector per pose)
    S* rowS = nullptr;             // [24][pc_chunks * pc_cp] per-row copies, component-major so that a warp's loads are contiguous:
                                   // 0-5 Hpp_ii, 6-11 M_i^-1, 12-17 / 18-23 the first two pose-pose blocks (symmetric by construction:
                                   // -J_s^T Omega J_s, so six values each)
    S* rS = nullptr;               // [3][pc_chunks * pc_cp] initial residual by chunk row (handed to the persistent kernel)
    S* xS = nullptr;               // [3][pc_chunks * pc_cp] solution by chunk row (updated in place every iteration)
    int Eb_pad = 0;
    double* scal = nullptr;  // [32] classic: rz, pAp, rz_new, rz0, done flag, iterations ...; fused: see solve_pcg.cu
    unsigned* bar = nullptr; // [4] grid barrier counter of the fused kernel
    int variant = 0;         // bos_options.pcg_variant: 0 = fused persistent kernel, 1 = classic multi-kernel loop
    // chain preconditioner of the fused kernel (bos_options.pcg_precond 0): per chunk, the block-tridiagonal matrix of the Schur
    // diagonal blocks and the pose-pose blocks between consecutive chunk rows (the odometry chain), solved exactly per CG iteration
    int precond = 0;         // 0 = chain + coarse space, 2 = chain only (both fall back to 1 when the factors do not fit in shared
                             // memory), 1 = 3x3 block-Jacobi
    S* chD = nullptr;        // [6][rows] Schur diagonal blocks (identity on padding rows)
    S* chO = nullptr;        // [6][rows] block between chunk row R and R + 1 (symmetric by construction), zero if none / across chunks
    float* chF = nullptr;    // [chunks][ch_fac_floats] factors in the shared-memory layout of the solve (see k_pcg_chain_factor)
    int ch_Kp = 0, ch_cps = 0, ch_fac_floats = 0;
    // coarse space of the chain preconditioner: piecewise-linear hats over the chunks (node c = start of chunk c, 3 dof per
    // node), Galerkin operator A_c = P^T S P assembled and inverted once per solve; z = M_chunk^-1 r + P A_c^-1 P^T r
    int c_nc = 0;            // 3 * (pc_chunks * c_nseg + 1)
    int c_h = 32, c_nseg = 1;   // node geometry: every chunk is cut into c_nseg segments of c_h rows (a multiple of 32), nodes at their ends
    int c_ld = 0;            // leading dimension of cAinv (c_nc rounded up to a multiple of 4: 32-byte aligned rows)
    int c_bw = 0;            // half bandwidth of A_c in scalars (from the pattern); 0 = treat A_c as dense (wide loop closures)
    double* cLc = nullptr;   // [c_nc][c_bw + 1] band of the Cholesky factor by columns, cLr the same by rows, cLdi = 1 / diagonal
    double* cLr = nullptr;
    double* cLdi = nullptr;
    // A_c^-1 is kept across GN steps (any SPD coarse operator is a valid preconditioner): refreshed every coarse_refresh solves, when the
    // state was replaced from outside (coarse_valid = false) and when the CG iteration count drifts (coarse_stale)
    bool coarse_valid = false, coarse_stale = false;
    int coarse_age = 0, coarse_refresh = 1, coarse_its_ref = 0;
    int coarse_period = 0, coarse_its_last = 0;   // (unused since the amortised schedule) / iterations of the previous solve
    // amortised refresh schedule: CG iterations spent above the post-refresh count since the last rebuild, and what a rebuild costs in CG iterations
    // (measured with CUDA events on every rebuilding solve: refresh kernels / (fused kernel / iterations))
    double coarse_excess = 0.0, coarse_ratio = 32.0;
    cudaEvent_t coarse_ev[3] = {nullptr, nullptr, nullptr};
    double* host_scal = nullptr;                  // 32 doubles of PINNED host memory (owned by the context): where the solve's scalars are read back
    double* cA = nullptr;    // [c_nc][c_nc] column-major lower: A_c, then its Cholesky factor
    double* cAinv = nullptr; // [c_nc][c_ld] A_c^-1 (full, symmetric)
    double* cRc = nullptr;   // [pc_chunks * c_nseg][6] per segment: P^T r restricted to its rows (left node, right node)
    double* cStats = nullptr;  // [8] scratch status of the coarse factorisation
    int sm_count = 148;
    int precond_used = 1;    // what the last solve actually ran (reported through bos_stats)
    int resolves = 0;        // 1: the last solve broke down under the chain / coarse preconditioner and was repeated with 3x3 blocks
};
template <typename S>
int launch_pcg_solve(const Dev<S>& d, PcgWork<S>& w, int max_iters, double rtol, cudaStream_t st,
                     int* iterations_out, int* launches);

template <typename S>
int dense_cholesky_lower(S* Smat, int n, double* stats, cudaStream_t st);

// Opt-in dynamic shared memory is a per-device attribute of a kernel: remember the largest size configured per
// (kernel, current device), so that contexts on several devices of one process all get it.
inline bool ensure_dyn_smem(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lk(mu);
    size_t& cur = done[std::make_pair(func, dev)];
    if (bytes <= cur) return true;
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return false;
    cur = bytes;
    return true;
}

// batched
template <typename S>
struct BatchDev {
    int nprob = 0, NP = 0, NL = 0, fixed = 0, Eb = 0, Eo = 0;
    const int* b_pose = nullptr;
    const int* b_lm = nullptr;
    const S* b_om = nullptr;   // [Eb]
    const S* b_z = nullptr;    // [nprob][Eb]
    const int* o_src = nullptr;
    const int* o_dst = nullptr;
    const S* o_om = nullptr;   // [Eo][6]
    const S* o_z = nullptr;    // [nprob][Eo][3]
    S* pose = nullptr;         // [nprob][NP][4]
    S* lm = nullptr;           // [nprob][NL][2]
    double* chi2 = nullptr;    // [nprob][2]
    double* delta_inf = nullptr;
    int* status = nullptr;
};
template <typename S>
int launch_batch_step(const BatchDev<S>& d, double kernel_threshold, double damping, cudaStream_t st);
size_t batch_smem_bytes(int NP, int NL, size_t scalar_bytes);

// ---- host-side pattern builder (pattern.cpp) --------------------------------------------------------
struct HostPattern {
    int NP = 0, NL = 0, fixed = 0, Eb = 0, Eo = 0, N = 0;
    // (pose, lm)-sorted bearing edges
    std::vector<int> b_pose, b_lm, b_perm, b_slot;
    bool slots_identity = true;
    std::vector<int> slot_pose, slot_lm;    // unique (pose, lm) blocks, sorted
    std::vector<int> pose_ptr, lm_ptr, lm_order, lm_order_pose, lm_order_lm;
    std::vector<int> o_src, o_dst, o_slot, oe_ptr, oe_edge, oe_other;
    std::vector<unsigned char> o_shared;
    bool has_shared_off = false;
    std::vector<int> off_lo, off_hi;        // unique pose-pose blocks, sorted
    std::vector<int> pp_ptr, pp_nbr, pp_slot;
    std::vector<int> tri_ptr, tri_edge;
    std::vector<int> pl_lm_id, b_row, ell_Loff, ell_Lmap, ell_Lpose;
    int pc_chunks = 0, pc_cp = 0;
    bool pc_ok = false;
    std::vector<int> pc_row_pose, pc_goff, pc_cl_ptr, pc_cl_row, pc_emap, pc_nbr, pc_nslot, pc_ncnt;
    std::vector<unsigned short> pc_loc;
    std::vector<int> lc_gptr, lc_goff, lc_emap, sh_ptr, sh_src, sh_ell;   // chunk-local landmark-major layout + landmark sharing lists (pattern.cpp)
    std::vector<unsigned short> lc_row, lc_k;
    std::vector<unsigned char> sh_first;
    std::vector<int> tile_ptr, tg_lm, tg_eptr, epose_ptr;
    std::vector<unsigned short> tg_edge;
    std::vector<char> touched;              // [NP + NL]
    // scalar CSC pattern of H_nofixed (slam/solver.cpp:72-75)
    bool csc_built = false;
    std::vector<int> csc_colptr, csc_rowidx;
    // where each CSC entry comes from: source kind (0 Hpp,1 Hll,2 Hoff,3 Hpl), flat index into the EXPANDED block arrays
    std::vector<int> csc_src_kind;
    std::vector<int64_t> csc_src_index;
    std::string error;
};
void build_csc(HostPattern& P);
// The bearing-edge core of the pattern (sorted edge order, block slots, CSR-of-blocks row pointers, landmark-major order, triangulation rows):
// built on the host by build_pattern's first phases, or on the device (setup.cu) and handed to build_pattern.
struct PatternCore {
    bool valid = false, slots_identity = true;
    std::vector<int> b_perm, b_pose, b_lm, b_slot, slot_pose, slot_lm, pose_ptr, lm_ptr, lm_order, lm_order_pose, lm_order_lm, tri_ptr, tri_edge, epose_ptr;
};
int build_pattern(HostPattern& P, int NP, int NL, int fixed, int64_t Eb, const int32_t* b_pose, const int32_t* b_lm,
                  int64_t Eo, const int32_t* o_src, const int32_t* o_dst, int pcg_chunks = 148, PatternCore* core = nullptr);
int device_pattern_core(PatternCore& core, int NP, int NL, int64_t Eb, const int32_t* b_pose, const int32_t* b_lm, cudaStream_t st, std::string& err);
int device_resolve_ids(int NP, const int32_t* pose_ids, int64_t Eb, const int32_t* b_pose_id, const int32_t* b_lm_id, int64_t Eo, const int32_t* o_src_id,
                       const int32_t* o_dst_id, int32_t* b_pose, int32_t* b_lm, int32_t* o_src, int32_t* o_dst, int32_t* lm_ids, int32_t* NL_out, cudaStream_t st,
                       std::string& err);

}  // namespace bos
