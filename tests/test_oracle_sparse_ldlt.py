"""The oracle's restatement of the reference's linear solve (Eigen::SimplicialLDLT: slam/solver.hpp:72, slam/solver.cpp:75-85):
minimum-degree ordering + symbolic phase once, up-looking LDL^T per step (oracle/bos_sparse_ldlt.hpp).  Pinned against the dense
LDL^T of the same oracle, against scipy's sparse LU on the exported CSC, and -- for the literal per-edge sparse merge of
slam/solver.cpp:44,60 -- against the O(E) block assembly."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from helpers import golden_problem, load_golden, oracle_for, synth_problem


def _nofixed(pr, v):
    keep = np.ones(len(v), bool)
    keep[3 * pr.fixed_stix:3 * pr.fixed_stix + 3] = False
    return v[keep]


@pytest.mark.parametrize("name", ["mini", "full"])
def test_sparse_ldlt_equals_dense_ldlt_on_the_bundled_datasets(name):
    g = load_golden(name)
    pr = golden_problem(g)
    o = oracle_for(g["pose_ids"], g["poses_xyt"], pr)
    o.linearize()
    o.solve(0); dense = o.delta()
    info = o.solve_sparse(); sparse = o.delta()
    assert info["finished"] and info["status"] == 0
    assert np.abs(dense - sparse).max() <= 1e-11 * np.abs(dense).max()
    assert np.all(sparse[3 * pr.fixed_stix:3 * pr.fixed_stix + 3] == 0)
    # the symbolic phase is cached (analyzePattern once, solver.cpp:77-80): a second call re-uses ordering and pattern
    o.step(2); o.linearize()
    info2 = o.solve_sparse()
    assert info2["t_order"] == info["t_order"] and info2["nnzL"] == info["nnzL"]
    # fill: the factor holds at least the lower triangle of H_nofixed, and the ordering keeps it well below dense
    colptr, rowidx, val, b = o.csc()
    n = len(b)
    lower = int(sum((rowidx[colptr[j]:colptr[j + 1]] > j).sum() for j in range(n)))
    assert lower <= info["nnzL"] <= n * (n - 1) // 2
    if name == "full":
        assert info["nnzL"] < 0.1 * n * n / 2


def test_sparse_ldlt_in_the_reference_precision():
    """FP32 (the reference's precision): the two factorisation orders agree to the rounding the conditioning of H allows (~1e6 on
    the bundled data), far looser than FP64 -- documented, not a parity claim."""
    g = load_golden("full")
    pr = golden_problem(g)
    o = oracle_for(g["pose_ids"], g["poses_xyt"], pr, "f32")
    o.linearize()
    o.solve(0); dense = o.delta()
    assert o.solve_sparse()["finished"]
    assert np.abs(dense - o.delta()).max() <= 5e-3 * np.abs(dense).max()


def test_sparse_ldlt_against_scipy_on_a_synthetic_world():
    w, pr = synth_problem(3000, 700, 30000, seed=5)
    o = oracle_for(w["pose_ids"], w["poses_init"], pr)
    o.linearize()
    info = o.solve_sparse()
    assert info["finished"] and info["status"] == 0
    colptr, rowidx, val, b = o.csc()
    n = len(b)
    H = sp.csc_matrix((val, rowidx, colptr), shape=(n, n))
    x = spl.splu(H).solve(-b)
    d = _nofixed(pr, o.delta())
    assert np.abs(d - x).max() <= 1e-9 * np.abs(x).max()
    assert np.abs(H @ d + b).max() <= 1e-11 * np.abs(b).max()
    # the minimum-degree ordering pays: at most a few times the entries of H itself
    assert info["nnzL"] < 6 * (len(val) - n) / 2
    # the same step through the three solver kinds of the oracle
    P, L = o.state()
    outs = []
    for kind in (2, 1):
        o.set_state(P, L)
        o.step(kind, 20000, 1e-13)
        outs.append(o.state())
    assert np.abs(outs[0][0] - outs[1][0]).max() <= 1e-8 and np.abs(outs[0][1] - outs[1][1]).max() <= 1e-8


def test_a_deadline_stops_the_factorisation():
    w, pr = synth_problem(3000, 700, 30000, seed=5)
    o = oracle_for(w["pose_ids"], w["poses_init"], pr)
    o.linearize()
    assert o.solve_sparse(1e-9)["finished"] is False


@pytest.mark.parametrize("name,dtype", [("mini", "f64"), ("full", "f64"), ("full", "f32")])
def test_literal_per_edge_merge_gives_the_block_assembly(name, dtype):
    """slam/solver.cpp:44,60 merges an N x N sparse temporary into H for every edge; same H and b as the O(E) assembly."""
    g = load_golden(name)
    pr = golden_problem(g)
    o = oracle_for(g["pose_ids"], g["poses_xyt"], pr, dtype)
    o.linearize()
    scale = np.abs(o.csc()[2]).max()
    assert o.literal_max_diff() <= (1e-12 if dtype == "f64" else 1e-5) * scale
    if name == "full":   # the literal accumulation really is O(N + nnz) per edge: orders of magnitude slower
        assert o.time_linearize_literal(1) > 20 * o.time_linearize(3)


def test_wrap_branch_hook_touches_only_branch_cut_edges():
    g = load_golden("full")
    pr = golden_problem(g)
    o = oracle_for(g["pose_ids"], g["poses_xyt"], pr)
    o.linearize()
    e0 = o.edge_terms()[0].copy(); b0 = o.blocks()["b"].copy()
    amb = np.where(np.abs(np.abs(e0) - np.pi) < 1e-9)[0]
    assert amb.tolist() == [29, 1324, 1515]                 # the single edges of landmarks 112, 114, 69 (slam/triangulation.cpp:38-42)
    o.set_wrap_branch(np.arange(pr.Eb), -np.sign(e0).astype(np.int32))   # ask to flip EVERY edge: only the three on the cut move
    o.linearize()
    e1 = o.edge_terms()[0]
    moved = np.where(e1 != e0)[0]
    assert moved.tolist() == amb.tolist() and np.allclose(e1[amb], -e0[amb])
    assert not np.array_equal(o.blocks()["b"], b0)
    o.set_wrap_branch(amb, np.sign(e0[amb]).astype(np.int32))
    o.linearize()
    assert np.array_equal(o.edge_terms()[0], e0) and np.array_equal(o.blocks()["b"], b0)
