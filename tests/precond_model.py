"""numpy / scipy model of the PCG preconditioners of csrc/solve_pcg.cu (test infrastructure, CPU only).

The GPU applies, per chunk of cp consecutive poses, the exact inverse of the block-tridiagonal matrix made of the Schur diagonal
blocks and the pose-pose blocks of consecutive poses (`k_pcg_chain_factor`, `chain_apply`), plus a Galerkin coarse space of
piecewise-linear hats over the chunks (`k_coarse_lm`, `k_coarse_pose`).  This file restates that operator with sparse LU / dense
inverses on the oracle's matrices, so that tests can check (a) that it is what the design says (SPD, same solution), and (b) that
the CUDA kernel needs the SAME number of CG iterations as the model -- a much sharper check of the factorisation, the two-level
solve and the coarse assembly than "it converges"."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

COARSE_MAX_CHUNKS = 8       # kCoarseMaxChunks in solve_pcg.cu (counts segments: the nodes sit at segment ends)


def coarse_geometry(cp, nodes_per_chunk=4):
    """ctx.cu ensure_pcg: every chunk is cut into nseg segments of h = 32 m rows (whole 32-row groups, at most 8 per chunk)"""
    groups = max(cp // 32, 1)
    want = min(min(nodes_per_chunk, 8), groups)
    m = (groups + want - 1) // want
    return 32 * m, (groups + m - 1) // m


def chunking(NP, sm_count=148):
    """pattern.cpp: chunk c owns poses [c * cp, (c + 1) * cp)"""
    nch = max(1, sm_count)
    nch = min(nch, (NP + 31) // 32)
    cp = ((NP + nch - 1) // nch + 31) // 32 * 32
    nch = (NP + cp - 1) // cp
    return nch, cp


def pose_system(o, NP, fixed_stix, damping=0.01):
    """Full-pose-space pieces of the oracle's linear system: Hpp (3NP x 3NP, the fixed pose a decoupled damping block), Hpl
    (3NP x 2NL, zero rows for the fixed pose), Hll^-1 (block diagonal), b_p, b_l."""
    colptr, rowidx, val, b = o.csc()
    n = len(b)
    A = sp.csc_matrix((val, rowidx, colptr), shape=(n, n)).tocsr()
    npz = 3 * (NP - 1)
    full = np.arange(npz)
    full[full >= 3 * fixed_stix] += 3
    E = sp.csr_matrix((np.ones(npz), (full, np.arange(npz))), shape=(3 * NP, npz))
    fx = np.arange(3 * fixed_stix, 3 * fixed_stix + 3)
    Hpp = (E @ A[:npz, :npz] @ E.T + sp.csr_matrix((np.full(3, damping), (fx, fx)), shape=(3 * NP, 3 * NP))).tocsr()
    Hpl = (E @ A[:npz, npz:]).tocsr()
    Hll = A[npz:, npz:].tobsr(blocksize=(2, 2))
    nl = Hll.shape[0] // 2
    assert Hll.nnz == 4 * nl
    Hlli = sp.bsr_matrix((np.linalg.inv(Hll.data), Hll.indices, Hll.indptr), shape=Hll.shape).tocsr()
    bp = E @ b[:npz]
    bl = b[npz:]
    return Hpp, Hpl, Hlli, bp, bl


def schur(Hpp, Hpl, Hlli, bp, bl):
    S = (Hpp - Hpl @ Hlli @ Hpl.T).tocsr()
    g = -(bp - Hpl @ (Hlli @ bl))
    return S, g


def block_diag3(S):
    Sb = S.tobsr(blocksize=(3, 3))
    nb = Sb.shape[0] // 3
    D = np.zeros((nb, 3, 3))
    rows = np.repeat(np.arange(nb), np.diff(Sb.indptr))
    m = rows == Sb.indices
    D[rows[m]] = Sb.data[m]
    return D


def block_jacobi(S):
    Dinv = np.linalg.inv(block_diag3(S))
    nb = Dinv.shape[0]
    M = sp.bsr_matrix((Dinv, np.arange(nb), np.arange(nb + 1)), shape=S.shape).tocsr()
    return lambda r: M @ r


def chain_blocks(S, Hpp, NP, cp):
    """block-tridiagonal chunk matrix: Schur diagonal blocks + the Hpp blocks between consecutive poses of one chunk"""
    nb = NP
    D = block_diag3(S)
    idx = np.arange(nb)
    off = np.ones(nb - 1)
    off[(idx[1:] % cp) == 0] = 0
    pat = sp.kron(sp.diags([off, off], [1, -1]), np.ones((3, 3))).tocsr()
    M = (sp.bsr_matrix((D, np.arange(nb), np.arange(nb + 1)), shape=S.shape) + Hpp.multiply(pat)).tocsc()
    lu = spl.splu(M, permc_spec="NATURAL")
    return M, lu.solve


def coarse_space(Hpp, Hpl, Hlli, NP, cp, nch, fixed_stix, h=None, nseg=1):
    """hats over the segments (node n = start of global segment n = chunk * nseg + j, one more node closes the last segment), Galerkin
    operator with the landmarks the kernel keeps (node list of at most 2 * COARSE_MAX_CHUNKS)"""
    if h is None:
        h = cp
    r = np.arange(NP) % cp
    c = np.arange(NP) // cp
    j = r // h
    seglen = np.minimum(h, cp - j * h)
    t = (r - j * h + 0.5) / seglen
    gs = c * nseg + j
    wl, wr = 1.0 - t, t.copy()
    wl[fixed_stix] = 0.0; wr[fixed_stix] = 0.0
    rows, cols, vals = [], [], []
    for a in range(3):
        rows += list(3 * np.arange(NP) + a); cols += list(3 * gs + a); vals += list(wl)
        rows += list(3 * np.arange(NP) + a); cols += list(3 * (gs + 1) + a); vals += list(wr)
    P = sp.csr_matrix((vals, (rows, cols)), shape=(3 * NP, 3 * (nch * nseg + 1)))
    # landmarks seen from too many segments are left out of the coarse operator
    Hb = Hpl.tocsc()
    nl = Hpl.shape[1] // 2
    keep = np.ones(2 * nl)
    for l in range(nl):
        col = Hb[:, 2 * l:2 * l + 2]
        poses = np.unique(col.nonzero()[0] // 3)
        poses = poses[poses != fixed_stix]
        # k_coarse_lm: segments ascending, neighbouring segments share a node, at most 2 * COARSE_MAX_CHUNKS nodes
        nn, last = 0, -2
        for sg in np.unique(gs[poses]):
            if nn + 2 > 2 * COARSE_MAX_CHUNKS:
                keep[2 * l:2 * l + 2] = 0.0
                break
            nn += 1 if last == sg else 2
            last = sg + 1
    K = sp.diags(keep)
    HP = Hpl.T @ P
    Ac = (P.T @ (Hpp @ P) - HP.T @ (K @ Hlli @ K) @ HP).toarray()
    dz = np.where(np.diag(Ac) <= 0)[0]
    Ac[dz, dz] = 1.0
    Aci = np.linalg.inv(Ac)
    return P, Ac, lambda rr: P @ (Aci @ (P.T @ rr))


def pcg(S, g, Minv, rtol, maxit=20000):
    """preconditioned CG with the kernel's stopping rule: r^T M^-1 r <= rtol^2 * its initial value"""
    x = np.zeros_like(g); r = g.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz; it = 0
    while it < maxit and rz > rtol * rtol * rz0:
        q = S @ p
        a = rz / (p @ q)
        x += a * p; r -= a * q
        z = Minv(r); rz2 = r @ z
        p = z + (rz2 / rz) * p; rz = rz2; it += 1
    return x, it
