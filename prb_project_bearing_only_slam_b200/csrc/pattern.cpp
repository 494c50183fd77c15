// pattern.cpp -- host-side, one-time, integer-only: the CSR-of-blocks sparsity pattern of H and the
// per-edge block slots.  Replaces what the reference rediscovers for every edge of every iteration
// (std::map id lookups, framework/state.cpp:43-63; sparse merges, slam/solver.cpp:44,60) and
// construct_the_permutation (slam/solver.cpp:99-125): the fixed pose's rows/cols are simply absent
// from the exported scalar pattern.  Everything here must be bit-exact against the oracle.
#include "bos_internal.h"

#include <algorithm>
#include <numeric>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace bos {

namespace {
// Static ranges of [0, n) on up to `threads` threads; fn(begin, end, thread index).  Every use below writes disjoint outputs per
// range, so the result does not depend on the thread count (tests/test_host_pattern.py compares checksums).
template <typename F>
void parallel_ranges(int n, int threads, F fn) {
    if (threads <= 1 || n < 2 * threads) { fn(0, n, 0); return; }
    std::vector<std::thread> pool;
    const int per = (n + threads - 1) / threads;
    for (int t = 1; t < threads; t++) {
        const int a = std::min(n, t * per), b = std::min(n, a + per);
        if (a < b) pool.emplace_back([=]() { fn(a, b, t); });
    }
    fn(0, std::min(n, per), 0);
    for (auto& th : pool) th.join();
}
int pattern_threads() {
    const char* env = std::getenv("BOS_PATTERN_THREADS");
    int t = env ? std::atoi(env) : (int)std::thread::hardware_concurrency();
    return t < 1 ? 1 : (t > 16 ? 16 : t);
}

struct PhaseTimer {   // BOS_PATTERN_TIMING=1: per-phase wall times of the pattern build on stderr
    bool on; std::chrono::steady_clock::time_point t0;
    PhaseTimer() : on(std::getenv("BOS_PATTERN_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[pattern] %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

int build_pattern(HostPattern& P, int NP, int NL, int fixed, int64_t Eb64, const int32_t* b_pose, const int32_t* b_lm,
                  int64_t Eo64, const int32_t* o_src, const int32_t* o_dst, int pcg_chunks, PatternCore* core) {
    P = HostPattern();
    if (NP <= 0 || NL < 0 || Eb64 < 0 || Eo64 < 0 || Eb64 > 0x3fffffff || Eo64 > 0x3fffffff) { P.error = "bad sizes"; return 1; }
    if (fixed < 0 || fixed >= NP) { P.error = "fixed pose stix out of range"; return 1; }
    const int Eb = (int)Eb64, Eo = (int)Eo64;
    P.NP = NP; P.NL = NL; P.fixed = fixed; P.Eb = Eb; P.Eo = Eo; P.N = 3 * NP + 2 * NL;
    for (int e = 0; e < Eb; e++)
        if (b_pose[e] < 0 || b_pose[e] >= NP || b_lm[e] < 0 || b_lm[e] >= NL) { P.error = "bearing edge index out of range"; return 1; }
    for (int e = 0; e < Eo; e++) {
        if (o_src[e] < 0 || o_src[e] >= NP || o_dst[e] < 0 || o_dst[e] >= NP) { P.error = "odometry edge index out of range"; return 1; }
        if (o_src[e] == o_dst[e]) { P.error = "odometry self-loop"; return 1; }
    }
    PhaseTimer tm;
    P.touched.assign((size_t)NP + NL, 0);
    tm.lap("checks");

    if (core && core->valid) {
        // the device built these tables (setup.cu): identical by construction, tests compare the checksum over ALL tables
        P.b_perm.swap(core->b_perm); P.b_pose.swap(core->b_pose); P.b_lm.swap(core->b_lm); P.b_slot.swap(core->b_slot);
        P.slot_pose.swap(core->slot_pose); P.slot_lm.swap(core->slot_lm); P.slots_identity = core->slots_identity;
        P.pose_ptr.swap(core->pose_ptr); P.lm_ptr.swap(core->lm_ptr); P.lm_order.swap(core->lm_order);
        P.lm_order_pose.swap(core->lm_order_pose); P.lm_order_lm.swap(core->lm_order_lm);
        P.tri_ptr.swap(core->tri_ptr); P.tri_edge.swap(core->tri_edge); P.epose_ptr.swap(core->epose_ptr);
        for (int i = 0; i < NP; i++) if (P.pose_ptr[i + 1] > P.pose_ptr[i]) P.touched[i] = 1;
        for (int j = 0; j < NL; j++) if (P.lm_ptr[j + 1] > P.lm_ptr[j]) P.touched[(size_t)NP + j] = 1;
        tm.lap("core tables from the device");
    } else {
    // ---- bearing edges sorted by (pose, lm), ties in caller order ------------------------------------------
    std::vector<uint64_t> key(Eb);
    P.b_perm.resize(Eb);
    std::iota(P.b_perm.begin(), P.b_perm.end(), 0);
    for (int e = 0; e < Eb; e++) key[e] = ((uint64_t)(uint32_t)b_pose[e] << 32) | (uint32_t)b_lm[e];
    bool sorted = true;
    for (int e = 1; e < Eb && sorted; e++) sorted = key[e - 1] <= key[e];
    if (!sorted)
        std::stable_sort(P.b_perm.begin(), P.b_perm.end(), [&](int a, int b) { return key[a] < key[b]; });
    P.b_pose.resize(Eb); P.b_lm.resize(Eb); P.b_slot.resize(Eb);
    P.slot_pose.clear(); P.slot_lm.clear();
    P.slots_identity = true;
    for (int k = 0; k < Eb; k++) {
        const int e = P.b_perm[k];
        P.b_pose[k] = b_pose[e]; P.b_lm[k] = b_lm[e];
        P.touched[b_pose[e]] = 1; P.touched[(size_t)NP + b_lm[e]] = 1;
        if (k == 0 || key[e] != key[P.b_perm[k - 1]]) { P.slot_pose.push_back(b_pose[e]); P.slot_lm.push_back(b_lm[e]); }
        P.b_slot[k] = (int)P.slot_pose.size() - 1;
        if (P.b_slot[k] != k) P.slots_identity = false;
    }
    tm.lap("sort + slots");
    const int n_hpl = (int)P.slot_pose.size();
    P.pose_ptr.assign(NP + 1, 0);
    for (int s = 0; s < n_hpl; s++) P.pose_ptr[P.slot_pose[s] + 1]++;
    for (int i = 0; i < NP; i++) P.pose_ptr[i + 1] += P.pose_ptr[i];
    // slots grouped by landmark, ascending pose inside a landmark (counting sort keeps slot order)
    P.lm_ptr.assign(NL + 1, 0);
    for (int s = 0; s < n_hpl; s++) P.lm_ptr[P.slot_lm[s] + 1]++;
    for (int j = 0; j < NL; j++) P.lm_ptr[j + 1] += P.lm_ptr[j];
    P.lm_order.resize(n_hpl); P.lm_order_pose.resize(n_hpl); P.lm_order_lm.resize(n_hpl);
    {
        std::vector<int> cur(P.lm_ptr.begin(), P.lm_ptr.end() - 1);
        for (int s = 0; s < n_hpl; s++) {
            int k = cur[P.slot_lm[s]]++;
            P.lm_order[k] = s; P.lm_order_pose[k] = P.slot_pose[s]; P.lm_order_lm[k] = P.slot_lm[s];
        }
    }
    tm.lap("pose/lm slot orders");
    // bearing edges grouped by landmark in caller order (triangulation rows, slam/triangulation.cpp:5-19)
    P.tri_ptr.assign(NL + 1, 0);
    for (int e = 0; e < Eb; e++) P.tri_ptr[b_lm[e] + 1]++;
    for (int j = 0; j < NL; j++) P.tri_ptr[j + 1] += P.tri_ptr[j];
    P.tri_edge.resize(Eb);
    {
        std::vector<int> inv(Eb);
        for (int k = 0; k < Eb; k++) inv[P.b_perm[k]] = k;
        std::vector<int> cur(P.tri_ptr.begin(), P.tri_ptr.end() - 1);
        for (int e = 0; e < Eb; e++) P.tri_edge[cur[b_lm[e]]++] = inv[e];
    }

    P.epose_ptr.assign(NP + 1, 0);
    for (int k = 0; k < Eb; k++) P.epose_ptr[P.b_pose[k] + 1]++;
    for (int i = 0; i < NP; i++) P.epose_ptr[i + 1] += P.epose_ptr[i];
    tm.lap("triangulation rows");
    }
    const int nthreads = pattern_threads();
    const bool par = nthreads > 1;
    const int wthreads = par ? std::max(1, (nthreads - 1) / 2) : 1;   // workers of each of the two big parts
    // tile-local grouping of the sorted bearing edges by landmark (static: depends only on the edge lists); independent of the
    // layouts below: runs on its own thread
    auto tile_grouping = [&]() {
        const int ntiles = (Eb + kLinTile - 1) / kLinTile;
        P.tile_ptr.assign(ntiles + 1, 0);
        P.tg_lm.clear(); P.tg_eptr.assign(1, 0); P.tg_edge.resize(Eb);
        struct Part { std::vector<int> lm, eptr, ngroups; };
        std::vector<Part> parts(wthreads > 0 ? wthreads : 1);
        parallel_ranges(ntiles, wthreads, [&](int t0, int t1, int w) {
            Part& Q = parts[w];
            std::vector<std::pair<int, int>> tmp;
            for (int t = t0; t < t1; t++) {
                const int a = t * kLinTile, b = std::min(Eb, a + kLinTile);
                tmp.clear();
                for (int k = a; k < b; k++) tmp.emplace_back(P.b_lm[k], k - a);
                std::sort(tmp.begin(), tmp.end());
                int ng = 0;
                for (size_t i = 0; i < tmp.size(); i++) {
                    if (i == 0 || tmp[i].first != tmp[i - 1].first) {
                        if (i) Q.eptr.push_back(a + (int)i);
                        Q.lm.push_back(tmp[i].first);
                        ng++;
                    }
                    P.tg_edge[a + i] = (unsigned short)tmp[i].second;
                }
                if (!tmp.empty()) Q.eptr.push_back(b);
                Q.ngroups.push_back(ng);
            }
        });
        int t = 0;
        for (const Part& Q : parts) {   // ranges are in tile order
            P.tg_lm.insert(P.tg_lm.end(), Q.lm.begin(), Q.lm.end());
            P.tg_eptr.insert(P.tg_eptr.end(), Q.eptr.begin(), Q.eptr.end());
            for (int ng : Q.ngroups) { P.tile_ptr[t + 1] = P.tile_ptr[t] + ng; t++; }
        }
    };
    // odometry edges and the pose-pose adjacency: independent of the bearing layouts, on its own thread
    auto odometry_tables = [&]() {
        // ---- odometry edges: unique unordered pose pairs ---------------------------------------------------------
        P.o_src.assign(o_src, o_src + Eo); P.o_dst.assign(o_dst, o_dst + Eo);
        std::vector<uint64_t> okey(Eo);
        for (int e = 0; e < Eo; e++) {
            int lo = std::min(o_src[e], o_dst[e]), hi = std::max(o_src[e], o_dst[e]);
            okey[e] = ((uint64_t)(uint32_t)lo << 32) | (uint32_t)hi;
            P.touched[o_src[e]] = 1; P.touched[o_dst[e]] = 1;
        }
        std::vector<uint64_t> uniq(okey);
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        const int n_off = (int)uniq.size();
        P.off_lo.resize(n_off); P.off_hi.resize(n_off);
        for (int k = 0; k < n_off; k++) { P.off_lo[k] = (int)(uniq[k] >> 32); P.off_hi[k] = (int)(uniq[k] & 0xffffffffu); }
        P.o_slot.resize(Eo);
        for (int e = 0; e < Eo; e++) P.o_slot[e] = (int)(std::lower_bound(uniq.begin(), uniq.end(), okey[e]) - uniq.begin());
        {
            std::vector<int> cnt(n_off, 0);
            for (int e = 0; e < Eo; e++) cnt[P.o_slot[e]]++;
            P.o_shared.resize(Eo);
            for (int e = 0; e < Eo; e++) { P.o_shared[e] = cnt[P.o_slot[e]] > 1; P.has_shared_off |= cnt[P.o_slot[e]] > 1; }
            // odometry edges incident to each pose (edge order inside a pose)
            P.oe_ptr.assign(NP + 1, 0);
            for (int e = 0; e < Eo; e++) { P.oe_ptr[o_src[e] + 1]++; P.oe_ptr[o_dst[e] + 1]++; }
            for (int i = 0; i < NP; i++) P.oe_ptr[i + 1] += P.oe_ptr[i];
            P.oe_edge.resize(2 * (size_t)Eo); P.oe_other.resize(2 * (size_t)Eo);
            std::vector<int> cur(P.oe_ptr.begin(), P.oe_ptr.end() - 1);
            for (int e = 0; e < Eo; e++) {
                int a = cur[o_src[e]]++; P.oe_edge[a] = (e << 1); P.oe_other[a] = o_dst[e];
                int b = cur[o_dst[e]]++; P.oe_edge[b] = (e << 1) | 1; P.oe_other[b] = o_src[e];
            }
        }
        // pose-pose adjacency, neighbours ascending
        P.pp_ptr.assign(NP + 1, 0);
        for (int k = 0; k < n_off; k++) { P.pp_ptr[P.off_lo[k] + 1]++; P.pp_ptr[P.off_hi[k] + 1]++; }
        for (int i = 0; i < NP; i++) P.pp_ptr[i + 1] += P.pp_ptr[i];
        P.pp_nbr.resize(2 * (size_t)n_off); P.pp_slot.resize(2 * (size_t)n_off);
        {
            std::vector<int> cur(P.pp_ptr.begin(), P.pp_ptr.end() - 1);
            // first the neighbours below a pose (it is the 'hi' side), in ascending lo; uniq is sorted by (lo, hi)
            for (int k = 0; k < n_off; k++) { int i = P.off_hi[k]; int c = cur[i]++; P.pp_nbr[c] = P.off_lo[k]; P.pp_slot[c] = k | (int)0x80000000; }
            for (int k = 0; k < n_off; k++) { int i = P.off_lo[k]; int c = cur[i]++; P.pp_nbr[c] = P.off_hi[k]; P.pp_slot[c] = k; }
        }

    };
    // BOS_PATTERN_THREADS=1 keeps everything on the calling thread (the results are identical: the three parts write disjoint
    // members of P)
    std::thread thB, thC;
    if (par) { thB = std::thread(tile_grouping); thC = std::thread(odometry_tables); }
    else { tile_grouping(); odometry_tables(); }
    // sliced-ELL layouts of the bearing EDGES for the fused PCG kernel (both coalesced for one-row-per-lane loops):
    //   L: rows = observed landmarks, renumbered compactly by descending observation count (uniform row length inside a
    //      group), kEllLanesL lanes per row (a row's edges go round-robin over its lanes), 32 / kEllLanesL rows per group;
    //   P: rows = poses in stix order, one lane per row, 32 rows per group.
    // A group is stored column-major: slot = (group offset + t) * 32 + lane; width = the group's longest row.
    {
        std::vector<int> eptr(NL + 1, 0), eord(Eb);
        for (int k = 0; k < Eb; k++) eptr[P.b_lm[k] + 1]++;
        for (int j = 0; j < NL; j++) eptr[j + 1] += eptr[j];
        {
            std::vector<int> cur(eptr.begin(), eptr.end() - 1);
            for (int k = 0; k < Eb; k++) eord[cur[P.b_lm[k]]++] = k;   // ascending sorted-edge index = ascending pose
        }
        P.pl_lm_id.clear();
        for (int l = 0; l < NL; l++)
            if (eptr[l + 1] > eptr[l]) P.pl_lm_id.push_back(l);
        std::stable_sort(P.pl_lm_id.begin(), P.pl_lm_id.end(),
                         [&](int a, int b) { return eptr[a + 1] - eptr[a] > eptr[b + 1] - eptr[b]; });
        const int n_clm = (int)P.pl_lm_id.size();
        P.b_row.assign(Eb, 0);
        constexpr int RPG = 32 / kEllLanesL;
        const int nLg = (n_clm + RPG - 1) / RPG;
        P.ell_Loff.assign(nLg + 1, 0);
        for (int g = 0; g < nLg; g++) {
            const int l0 = P.pl_lm_id[g * RPG];
            P.ell_Loff[g + 1] = P.ell_Loff[g] + (eptr[l0 + 1] - eptr[l0] + kEllLanesL - 1) / kEllLanesL;
        }
        P.ell_Lmap.assign((size_t)P.ell_Loff[nLg] * 32, -1);
        P.ell_Lpose.assign((size_t)P.ell_Loff[nLg] * 32, -1);
        parallel_ranges(n_clm, wthreads, [&](int r0, int r1, int) {
            for (int r = r0; r < r1; r++) {
                const int l = P.pl_lm_id[r], g = r / RPG, lane0 = (r % RPG) * kEllLanesL;
                for (int q = eptr[l]; q < eptr[l + 1]; q++) {
                    const int k = eord[q], idx = q - eptr[l];
                    const size_t slot = ((size_t)P.ell_Loff[g] + idx / kEllLanesL) * 32 + lane0 + idx % kEllLanesL;
                    P.ell_Lmap[slot] = k; P.ell_Lpose[slot] = (P.b_pose[k] == fixed) ? -1 : P.b_pose[k];
                    P.b_row[k] = r;
                }
            }
        });
        // P: pose rows in CHUNKS: chunk c (one persistent CTA) owns poses [c * cp, (c + 1) * cp), rows in pose order; every
        // slot names its landmark by a 16-bit index into the chunk's table of distinct landmark rows (a contiguous pose
        // range sees few landmarks).
        int nch = pcg_chunks < 1 ? 1 : pcg_chunks;
        if (nch > (NP + 31) / 32) nch = (NP + 31) / 32;
        const int cp = ((NP + nch - 1) / nch + 31) / 32 * 32;
        nch = (NP + cp - 1) / cp;
        P.pc_chunks = nch; P.pc_cp = cp;
        const int gpc = cp / 32;
        P.pc_row_pose.assign((size_t)nch * cp, -1);
        P.pc_goff.assign((size_t)nch * gpc + 1, 0);
        P.pc_cl_ptr.assign(nch + 1, 0);
        P.pc_cl_row.clear();
        P.pc_ok = true;
        std::vector<int> wdt((size_t)nch * gpc, 0);
        std::vector<std::vector<int>> cl(nch);
        parallel_ranges(nch, wthreads, [&](int c0, int c1, int) {
            for (int c = c0; c < c1; c++) {
                const int p0 = c * cp, p1 = std::min(NP, p0 + cp);
                for (int i = p0; i < p1; i++) {
                    const int r = i - p0;
                    P.pc_row_pose[(size_t)c * cp + r] = i;
                    int& wd = wdt[(size_t)c * gpc + r / 32];   // rows stay in pose order: the group is as wide as its longest row
                    wd = std::max(wd, P.epose_ptr[i + 1] - P.epose_ptr[i]);
                }
                // distinct landmark rows of the chunk, ascending
                std::vector<int>& v = cl[c];
                v.assign(P.b_row.begin() + P.epose_ptr[p0], P.b_row.begin() + P.epose_ptr[p1]);
                std::sort(v.begin(), v.end());
                v.erase(std::unique(v.begin(), v.end()), v.end());
            }
        });
        for (size_t q = 0; q < wdt.size(); q++) P.pc_goff[q + 1] = P.pc_goff[q] + wdt[q];
        for (int c = 0; c < nch; c++) {
            P.pc_cl_row.insert(P.pc_cl_row.end(), cl[c].begin(), cl[c].end());
            P.pc_cl_ptr[c + 1] = (int)P.pc_cl_row.size();
            if (cl[c].size() >= 0xffff) P.pc_ok = false;
        }
        P.pc_loc.assign((size_t)P.pc_goff.back() * 32, (unsigned short)0xffff);
        P.pc_emap.assign((size_t)P.pc_goff.back() * 32, -1);
        if (P.pc_ok)
            parallel_ranges(nch, wthreads, [&](int c0, int c1, int) {
                std::vector<int> lid(n_clm > 0 ? n_clm : 1, -1);   // landmark row -> index in the chunk's table
                for (int c = c0; c < c1; c++) {
                    const int cl0 = P.pc_cl_ptr[c], cl1 = P.pc_cl_ptr[c + 1];
                    for (int q = cl0; q < cl1; q++) lid[P.pc_cl_row[q]] = q - cl0;
                    for (int r = 0; r < cp; r++) {
                        const int i = P.pc_row_pose[(size_t)c * cp + r];
                        if (i < 0) continue;
                        const size_t g = (size_t)c * gpc + r / 32;
                        for (int k = P.epose_ptr[i]; k < P.epose_ptr[i + 1]; k++) {
                            const size_t slot = ((size_t)P.pc_goff[g] + (k - P.epose_ptr[i])) * 32 + r % 32;
                            P.pc_loc[slot] = (unsigned short)lid[P.b_row[k]];
                            P.pc_emap[slot] = k;
                        }
                    }
                }
            });
        // LC: per chunk, the landmark-major view of the chunk's OWN edges: rows = the chunk's distinct landmarks in table order,
        // kLcLanes (= 1) lanes per row (a landmark has ~10 edges inside a chunk, the busiest ~20; measured at synth-2M: two lanes per row
        // change nothing, four cost 5 %), 32 / kLcLanes rows per group in descending edge count (the kernel hands the groups to its warps
        // longest first), column-major groups like the L layout.  The persistent PCG kernel
        // forms, per chunk, the partial t_l = sum_k jh_k (Jp_k . z) of every landmark it sees from SHARED MEMORY (its poses' z and
        // positions live there) instead of gathering 48 bytes per edge from global memory; a landmark seen from several chunks is
        // summed from the chunks' partials through the sharing lists below.
        //   lc_gptr[c]   first group of chunk c (groups numbered globally)       lc_goff[g]  column offset of group g
        //   lc_row[slot] chunk-local pose row (0xffff: padding, or an edge of the fixed pose)   lc_emap[slot] sorted edge (-1: none)
        //   q = pc_cl_ptr[c] + k numbers every (chunk, local landmark): sh_ptr / sh_src list the q' of ALL chunks (own included) that
        //   see the same landmark, sh_first[q] = 1 in the lowest such chunk (it accounts for the landmark's part of z . S z)
        if (P.pc_ok) {
            constexpr int RPGc = 32 / kLcLanes;
            P.lc_gptr.assign(nch + 1, 0);
            for (int c = 0; c < nch; c++) P.lc_gptr[c + 1] = P.lc_gptr[c] + (P.pc_cl_ptr[c + 1] - P.pc_cl_ptr[c] + RPGc - 1) / RPGc;
            const int ngl = P.lc_gptr[nch];
            std::vector<int> gw(ngl > 0 ? ngl : 1, 0);
            P.lc_k.assign((size_t)(ngl > 0 ? ngl : 1) * 32, (unsigned short)0xffff);
            std::vector<std::vector<std::pair<unsigned short, int>>> ledges(P.pc_cl_row.size());   // per q: (local row, sorted edge)
            parallel_ranges(nch, wthreads, [&](int c0, int c1, int) {
                std::vector<int> lid(n_clm > 0 ? n_clm : 1, -1);
                for (int c = c0; c < c1; c++) {
                    const int cl0 = P.pc_cl_ptr[c], cl1 = P.pc_cl_ptr[c + 1];
                    for (int q = cl0; q < cl1; q++) lid[P.pc_cl_row[q]] = q - cl0;
                    for (int r = 0; r < cp; r++) {
                        const int i = P.pc_row_pose[(size_t)c * cp + r];
                        if (i < 0 || i == fixed) continue;
                        for (int k = P.epose_ptr[i]; k < P.epose_ptr[i + 1]; k++) ledges[cl0 + lid[P.b_row[k]]].emplace_back((unsigned short)r, k);
                    }
                    // ELL rows in descending edge count (uniform groups); lc_k names the local landmark of every row
                    std::vector<int> ord(cl1 - cl0);
                    for (int q = cl0; q < cl1; q++) ord[q - cl0] = q;
                    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return ledges[a].size() > ledges[b].size(); });
                    for (int j = 0; j < cl1 - cl0; j++) {
                        P.lc_k[(size_t)P.lc_gptr[c] * 32 + j] = (unsigned short)(ord[j] - cl0);
                        int& wd = gw[P.lc_gptr[c] + j / RPGc];
                        wd = std::max(wd, ((int)ledges[ord[j]].size() + kLcLanes - 1) / kLcLanes);
                    }
                }
            });
            P.lc_goff.assign(ngl + 1, 0);
            for (int g = 0; g < ngl; g++) P.lc_goff[g + 1] = P.lc_goff[g] + gw[g];
            P.lc_row.assign((size_t)P.lc_goff[ngl] * 32, (unsigned short)0xffff);
            P.lc_emap.assign((size_t)P.lc_goff[ngl] * 32, -1);
            parallel_ranges(nch, wthreads, [&](int c0, int c1, int) {
                for (int c = c0; c < c1; c++) {
                    const int cl0 = P.pc_cl_ptr[c], cl1 = P.pc_cl_ptr[c + 1];
                    for (int j = 0; j < cl1 - cl0; j++) {
                        const int q = cl0 + P.lc_k[(size_t)P.lc_gptr[c] * 32 + j];
                        const int g = P.lc_gptr[c] + j / RPGc, lane0 = (j % RPGc) * kLcLanes;
                        for (size_t idx = 0; idx < ledges[q].size(); idx++) {
                            const size_t slot = ((size_t)P.lc_goff[g] + idx / kLcLanes) * 32 + lane0 + idx % kLcLanes;
                            P.lc_row[slot] = ledges[q][idx].first;
                            P.lc_emap[slot] = ledges[q][idx].second;
                        }
                    }
                }
            });
            // sharing lists
            std::vector<std::vector<int>> seen(n_clm > 0 ? n_clm : 1);
            for (size_t q = 0; q < P.pc_cl_row.size(); q++) seen[P.pc_cl_row[q]].push_back((int)q);   // ascending q = ascending chunk
            P.sh_ptr.assign(P.pc_cl_row.size() + 1, 0);
            for (size_t q = 0; q < P.pc_cl_row.size(); q++) P.sh_ptr[q + 1] = P.sh_ptr[q] + (int)seen[P.pc_cl_row[q]].size();
            P.sh_src.resize((size_t)P.sh_ptr.back());
            P.sh_first.assign(P.pc_cl_row.size(), 0);
            for (size_t q = 0; q < P.pc_cl_row.size(); q++) {
                const std::vector<int>& v = seen[P.pc_cl_row[q]];
                std::copy(v.begin(), v.end(), P.sh_src.begin() + P.sh_ptr[q]);
                P.sh_first[q] = (v[0] == (int)q) ? 1 : 0;
            }
            // the first kShareEll sharers of every q once more, transposed ([j][n_q], -1 = none): consecutive threads read consecutive
            // words and all of a landmark's sources are known after ONE round trip (the CSR above serves the rare longer lists)
            const size_t nq = P.pc_cl_row.size();
            P.sh_ell.assign((size_t)kShareEll * (nq > 0 ? nq : 1), -1);
            for (size_t q = 0; q < nq; q++)
                for (int j = 0; j < kShareEll && P.sh_ptr[q] + j < P.sh_ptr[q + 1]; j++) P.sh_ell[(size_t)j * nq + q] = P.sh_src[P.sh_ptr[q] + j];
        }
        // up to two pose-pose neighbours per row inline (filled once the adjacency exists, below)
        P.pc_nbr.assign((size_t)nch * cp * 2, -1);
        P.pc_nslot.assign((size_t)nch * cp * 2, 0);
    }
    tm.lap("ELL + chunk layouts");
    if (par) thB.join();
    tm.lap("tile grouping");
    if (par) thC.join();
    // inline pose-pose neighbours of the PCG chunk rows (needs the adjacency above)
    P.pc_ncnt.assign((size_t)P.pc_chunks * P.pc_cp, 0);
    for (size_t R = 0; R < P.pc_row_pose.size(); R++) {
        const int i = P.pc_row_pose[R];
        if (i < 0) continue;
        const int q0 = P.pp_ptr[i], q1 = P.pp_ptr[i + 1];
        P.pc_ncnt[R] = q1 - q0;
        const size_t nrows = P.pc_row_pose.size();
        for (int k = 0; k < 2 && q0 + k < q1; k++) { P.pc_nbr[k * nrows + R] = P.pp_nbr[q0 + k]; P.pc_nslot[k * nrows + R] = P.pp_slot[q0 + k]; }
    }

    tm.lap("odometry + adjacency");
    P.csc_built = false;   // the scalar CSC view is only needed for inspection / parity downloads: built on demand (build_csc)
    return 0;
}

// Scalar CSC pattern of H_nofixed (slam/solver.cpp:72-75) with the source of every entry.  Only the inspection entry points
// (bos_download_csc, bos_pattern_info_get, bos_host_pattern_*) need it, so it is built lazily: at 2 M edges it is 27 M entries.
void build_csc(HostPattern& P) {
    if (P.csc_built) return;
    const int NP = P.NP, NL = P.NL, fixed = P.fixed;
    auto nofixed = [&](int i) { const int f3 = 3 * fixed; return i < f3 ? i : (i < f3 + 3 ? -1 : i - 3); };
    const int n = P.N - 3;
    P.csc_colptr.assign(n + 1, 0);
    P.csc_rowidx.clear(); P.csc_src_kind.clear(); P.csc_src_index.clear();
    auto emit = [&](int grow, int kind, int64_t idx) {
        int r = nofixed(grow);
        if (r < 0) return;
        P.csc_rowidx.push_back(r); P.csc_src_kind.push_back(kind); P.csc_src_index.push_back(idx);
    };
    for (int p = 0; p < NP; p++) {
        if (p == fixed) continue;
        for (int c = 0; c < 3; c++) {
            // neighbours below p, p itself, neighbours above p, then landmarks
            int a0 = P.pp_ptr[p], a1 = P.pp_ptr[p + 1];
            int q = a0;
            for (; q < a1 && P.pp_nbr[q] < p; q++) {
                int slot = P.pp_slot[q] & 0x7fffffff;  // block H[nbr][p]: row a in nbr, col c in p
                for (int a = 0; a < 3; a++) emit(3 * P.pp_nbr[q] + a, 2, (int64_t)slot * 9 + a * 3 + c);
            }
            for (int a = 0; a < 3; a++)
                if (P.touched[p] || a == c) emit(3 * p + a, 0, (int64_t)p * 9 + a * 3 + c);
            for (; q < a1; q++) {
                int slot = P.pp_slot[q] & 0x7fffffff;  // block H[p][nbr]: entry (row a in nbr, col c in p) is its transpose [c][a]
                for (int a = 0; a < 3; a++) emit(3 * P.pp_nbr[q] + a, 2, (int64_t)slot * 9 + c * 3 + a);
            }
            for (int s = P.pose_ptr[p]; s < P.pose_ptr[p + 1]; s++)
                for (int a = 0; a < 2; a++) emit(3 * NP + 2 * P.slot_lm[s] + a, 3, (int64_t)s * 6 + c * 2 + a);
            P.csc_colptr[nofixed(3 * p + c) + 1] = (int)P.csc_rowidx.size();
        }
    }
    for (int l = 0; l < NL; l++) {
        for (int c = 0; c < 2; c++) {
            for (int k = P.lm_ptr[l]; k < P.lm_ptr[l + 1]; k++) {
                int s = P.lm_order[k];
                for (int a = 0; a < 3; a++) emit(3 * P.slot_pose[s] + a, 3, (int64_t)s * 6 + a * 2 + c);
            }
            for (int a = 0; a < 2; a++)
                if (P.touched[(size_t)NP + l] || a == c) emit(3 * NP + 2 * l + a, 1, (int64_t)l * 4 + a * 2 + c);
            P.csc_colptr[nofixed(3 * NP + 2 * l + c) + 1] = (int)P.csc_rowidx.size();
        }
    }
    P.csc_built = true;
}

}  // namespace bos
