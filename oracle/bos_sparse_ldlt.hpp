// bos_sparse_ldlt.hpp -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see bos_oracle.hpp).
//
// Restatement of what the reference's linear solve does (slam/solver.hpp:72, slam/solver.cpp:75-85):
//     Eigen::SimplicialLDLT<SparseMatrix<float>> solver;  analyzePattern(H_nofixed) ONCE, then per step
//     factorize(H_nofixed) and solve(-b_nofixed).
// Eigen (un-vendored dependency, version unpinned, >= 3.3; absent from this image) implements SimplicialLDLT as
//   * a fill-reducing ordering: AMD (approximate minimum degree, Amestoy / Davis / Duff) on the pattern of A;
//   * analyzePattern: elimination tree + column counts of L by the row-subtree traversal of T. Davis' LDL package;
//   * factorize: the up-looking LDL^T of the same package, one sparse triangular solve per row, no pivoting, no supernodes,
//     single thread;
//   * solve: P^T L^-T D^-1 L^-1 P b.
// This file restates those published algorithms: `min_degree_order` is a quotient-graph minimum-degree ordering with
// element absorption and the AMD approximate external degree, run on the BLOCK graph (every pose / landmark block is a
// supervariable of weight 3 / 2, which is what AMD's indistinguishable-node detection finds on this matrix anyway);
// `SparseLdlt` is LDL's symbolic + numeric phases and the three solve sweeps.  The ordering only changes rounding and fill,
// never the mathematics; parity against the dense LDL^T of bos_oracle.hpp and scipy is tested in
// tests/test_oracle_sparse_ldlt.py.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <vector>

namespace bos_oracle {

// Quotient-graph minimum degree on a symmetric graph given as CSR adjacency (no self loops needed; duplicates tolerated).
// weight[i] = number of scalar rows of node i.  Returns the elimination order (node indices).
inline std::vector<int> min_degree_order(int n, const std::vector<int>& xadj, const std::vector<int>& adjncy, const std::vector<int>& weight) {
    std::vector<std::vector<int>> adj(n), elems(n), Le(n);
    std::vector<int> Lw(n, 0);                     // weight of an element's variable list
    std::vector<char> state(n, 0);                 // 0 variable, 1 element (eliminated, alive), 2 absorbed element
    std::vector<long long> deg(n, 0);
    long long total_w = 0;
    for (int i = 0; i < n; i++) {
        total_w += weight[i];
        adj[i].assign(adjncy.begin() + xadj[i], adjncy.begin() + xadj[i + 1]);
        std::sort(adj[i].begin(), adj[i].end());
        adj[i].erase(std::unique(adj[i].begin(), adj[i].end()), adj[i].end());
        adj[i].erase(std::remove(adj[i].begin(), adj[i].end(), i), adj[i].end());
    }
    for (int i = 0; i < n; i++)
        for (int v : adj[i]) deg[i] += weight[v];
    // degree buckets (doubly linked lists)
    const long long maxdeg = total_w + 1;
    std::vector<int> head((size_t)maxdeg + 1, -1), nxt(n, -1), prv(n, -1);
    auto bucket_insert = [&](int i) {
        const long long d = std::min(deg[i], maxdeg);
        nxt[i] = head[d]; prv[i] = -1;
        if (head[d] >= 0) prv[head[d]] = i;
        head[d] = i;
    };
    auto bucket_remove = [&](int i) {
        const long long d = std::min(deg[i], maxdeg);
        if (prv[i] >= 0) nxt[prv[i]] = nxt[i]; else head[d] = nxt[i];
        if (nxt[i] >= 0) prv[nxt[i]] = prv[i];
    };
    for (int i = 0; i < n; i++) bucket_insert(i);
    std::vector<int> mark(n, -1), wstamp(n, -1), order;
    std::vector<long long> w(n, 0);
    order.reserve(n);
    long long mindeg = 0, remaining_w = total_w;
    std::vector<int> Lp;
    for (int step = 0; step < n; step++) {
        while (mindeg <= maxdeg && head[mindeg] < 0) mindeg++;
        const int p = head[mindeg];
        bucket_remove(p);
        // ---- the new element: variables adjacent to p directly or through its elements --------------------------------
        Lp.clear();
        mark[p] = step;
        long long lpw = 0;
        for (int v : adj[p])
            if (state[v] == 0 && mark[v] != step) { mark[v] = step; Lp.push_back(v); lpw += weight[v]; }
        for (int e : elems[p]) {
            if (state[e] != 1) continue;
            for (int v : Le[e])
                if (state[v] == 0 && mark[v] != step) { mark[v] = step; Lp.push_back(v); lpw += weight[v]; }
            state[e] = 2;                           // absorbed into p
            std::vector<int>().swap(Le[e]);
        }
        std::vector<int>().swap(adj[p]);
        std::vector<int>().swap(elems[p]);
        state[p] = 1;
        remaining_w -= weight[p];
        // ---- prune the lists of the element's variables; w[e] = |L_e \ L_p| for every element they still touch -----------
        for (int i : Lp) {
            bucket_remove(i);
            auto& a = adj[i];
            size_t k = 0;
            for (int v : a)
                if (state[v] == 0 && mark[v] != step) a[k++] = v;
            a.resize(k);
            auto& el = elems[i];
            k = 0;
            for (int e : el)
                if (state[e] == 1 && e != p) {
                    el[k++] = e;
                    if (wstamp[e] != step) { wstamp[e] = step; w[e] = Lw[e]; }
                    w[e] -= weight[i];
                }
            el.resize(k);
        }
        // ---- approximate external degrees (AMD): d_i = |A_i \ L_p| + |L_p \ i| + sum_e |L_e \ L_p|, bounded --------------
        for (int i : Lp) {
            long long d = lpw - weight[i];
            for (int v : adj[i]) d += weight[v];
            auto& el = elems[i];
            size_t k = 0;
            for (int e : el) {
                if (w[e] <= 0) { state[e] = 2; std::vector<int>().swap(Le[e]); continue; }   // aggressive absorption: L_e is inside L_p
                el[k++] = e;
                d += w[e];
            }
            el.resize(k);
            el.push_back(p);
            d = std::min(d, remaining_w - weight[i]);
            d = std::min(d, deg[i] + lpw - weight[i]);
            deg[i] = d < 0 ? 0 : d;
            bucket_insert(i);
            if (deg[i] < mindeg) mindeg = deg[i];
        }
        Le[p] = Lp;
        Lw[p] = (int)lpw;
        order.push_back(p);
    }
    // absorbed elements keep Le empty; nothing else to clean
    return order;
}

// LDL^T of a symmetric matrix given as full CSC (both triangles, sorted or unsorted rows), with a fixed symmetric permutation.
template <class T>
struct SparseLdlt {
    int n = 0;
    std::vector<int> perm, pinv;        // perm[k] = original index of permuted row k
    std::vector<int> parent, Lp, Lnz, Li;
    std::vector<T> Lx, D;
    bool analyzed = false;
    int status = 0;                     // 1: a non-positive pivot was met (the reference only prints a message, solver.cpp:82-84)
    long long nnzL = 0;
    double flops = 0;                   // of one numeric factorisation: sum over columns of c (c + 2), c = entries of the column of L
    double flops_done = 0;              // same count for the rows the last factorize() call completed (== flops when it finished)
    int rows_done = 0;

    // analyzePattern (solver.cpp:77-80): elimination tree + column counts for the permuted matrix
    void analyze(int n_, const std::vector<int>& colptr, const std::vector<int>& rowidx, const std::vector<int>& perm_) {
        n = n_;
        perm = perm_;
        pinv.assign(n, 0);
        for (int k = 0; k < n; k++) pinv[perm[k]] = k;
        parent.assign(n, -1);
        Lnz.assign(n, 0);
        std::vector<int> flag(n);
        for (int k = 0; k < n; k++) {
            flag[k] = k;
            const int c = perm[k];
            for (int q = colptr[c]; q < colptr[c + 1]; q++) {
                int i = pinv[rowidx[q]];
                if (i < k)
                    for (; flag[i] != k; i = parent[i]) {
                        if (parent[i] == -1) parent[i] = k;
                        Lnz[i]++;
                        flag[i] = k;
                    }
            }
        }
        Lp.assign(n + 1, 0);
        flops = 0;
        for (int k = 0; k < n; k++) { Lp[k + 1] = Lp[k] + Lnz[k]; flops += (double)Lnz[k] * ((double)Lnz[k] + 2.0); }
        nnzL = Lp[n];
        Li.assign((size_t)nnzL, 0);
        Lx.assign((size_t)nnzL, T(0));
        D.assign(n, T(0));
        analyzed = true;
    }

    // factorize (solver.cpp:81): up-looking LDL^T.  deadline_s > 0 aborts (returns false) when the wall clock passes it.
    bool factorize(const std::vector<int>& colptr, const std::vector<int>& rowidx, const std::vector<T>& val, double deadline_s = 0.0) {
        std::vector<T> Y(n, T(0));
        std::vector<int> pattern(n), flag(n);
        std::fill(Lnz.begin(), Lnz.end(), 0);
        status = 0;
        flops_done = 0; rows_done = 0;
        double fl = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < n; k++) {
            if (deadline_s > 0.0 && (k & 1023) == 0 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > deadline_s) {
                flops_done = fl; rows_done = k;
                return false;
            }
            int top = n;
            flag[k] = k;
            const int c = perm[k];
            for (int q = colptr[c]; q < colptr[c + 1]; q++) {
                int i = pinv[rowidx[q]];
                if (i > k) continue;
                Y[i] += val[q];
                int len = 0;
                for (; flag[i] != k; i = parent[i]) { pattern[len++] = i; flag[i] = k; }
                while (len > 0) pattern[--top] = pattern[--len];
            }
            T dk = Y[k];
            Y[k] = T(0);
            for (; top < n; top++) {
                const int i = pattern[top];
                const T yi = Y[i];
                Y[i] = T(0);
                const int p2 = Lp[i] + Lnz[i];
                fl += 2.0 * Lnz[i] + 3.0;
                for (int q = Lp[i]; q < p2; q++) Y[Li[q]] -= Lx[q] * yi;
                const T lki = yi / D[i];
                dk -= lki * yi;
                Li[p2] = k;
                Lx[p2] = lki;
                Lnz[i]++;
            }
            D[k] = dk;
            if (!(dk > T(0))) status = 1;
        }
        flops_done = fl; rows_done = n;
        return true;
    }

    // solve (solver.cpp:85): x = P^T L^-T D^-1 L^-1 P rhs
    void solve(const std::vector<T>& rhs, std::vector<T>& x) const {
        std::vector<T> y(n);
        for (int k = 0; k < n; k++) y[k] = rhs[perm[k]];
        for (int j = 0; j < n; j++) {
            const T yj = y[j];
            for (int q = Lp[j]; q < Lp[j] + Lnz[j]; q++) y[Li[q]] -= Lx[q] * yj;
        }
        for (int j = 0; j < n; j++) y[j] /= D[j];
        for (int j = n - 1; j >= 0; j--) {
            T s = y[j];
            for (int q = Lp[j]; q < Lp[j] + Lnz[j]; q++) s -= Lx[q] * y[Li[q]];
            y[j] = s;
        }
        x.assign(n, T(0));
        for (int k = 0; k < n; k++) x[perm[k]] = y[k];
    }
};

}  // namespace bos_oracle
