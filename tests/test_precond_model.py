"""CPU model of the PCG preconditioners (tests/precond_model.py) on the oracle's matrices: the operator the CUDA kernel is
designed to apply is symmetric positive definite, gives the direct solution, and orders the iteration counts as DESIGN.md says
(3x3 block-Jacobi >> chunk chain blocks > chain blocks + coarse space).  The GPU test compares the kernel's counts with these."""
import numpy as np
import scipy.sparse.linalg as spl

from helpers import synth_problem, oracle_for
import precond_model as pm


def model_counts(NP, NL, E, seed, rtol, sm_count=148, nodes_per_chunk=4, with_bj=True):
    w, pr = synth_problem(NP, NL, E, seed=seed)
    o = oracle_for(w["pose_ids"], w["poses_init"], pr)
    o.linearize()
    Hpp, Hpl, Hlli, bp, bl = pm.pose_system(o, pr.NP, pr.fixed_stix)
    S, g = pm.schur(Hpp, Hpl, Hlli, bp, bl)
    nch, cp = pm.chunking(pr.NP, sm_count)
    M, chain = pm.chain_blocks(S, Hpp, pr.NP, cp)
    h, nseg = pm.coarse_geometry(cp, nodes_per_chunk)
    P, Ac, coarse = pm.coarse_space(Hpp, Hpl, Hlli, pr.NP, cp, nch, pr.fixed_stix, h, nseg)
    x_bj, it_bj = pm.pcg(S, g, pm.block_jacobi(S), rtol) if with_bj else (None, -1)
    x_ch, it_ch = pm.pcg(S, g, chain, rtol)
    x_cc, it_cc = pm.pcg(S, g, lambda r: chain(r) + coarse(r), rtol)
    return dict(S=S, g=g, M=M, Ac=Ac, x=(x_cc, x_bj, x_ch), its=(it_cc, it_bj, it_ch), w=w, pr=pr, o=o, nch=nch, cp=cp, h=h, nseg=nseg)


def test_chain_and_coarse_preconditioners_are_spd_and_cut_the_iterations():
    m = model_counts(3000, 700, 30000, seed=5, rtol=1e-10)
    S, g = m["S"], m["g"]
    x_ref = spl.spsolve(S.tocsc(), g)
    for x in m["x"]:
        assert np.abs(x - x_ref).max() <= 1e-7 * np.abs(x_ref).max()
    # the chunk matrix M = H_chain + blockdiag(bearing Schur diagonal + damping) and the Galerkin operator are SPD
    Md = m["M"].toarray()
    assert np.abs(Md - Md.T).max() <= 1e-9 * np.abs(Md).max() and np.linalg.eigvalsh(0.5 * (Md + Md.T)).min() > 0
    assert np.linalg.eigvalsh(0.5 * (m["Ac"] + m["Ac"].T)).min() > 0
    it_cc, it_bj, it_ch = m["its"]
    assert it_ch < 0.5 * it_bj and it_cc <= it_ch
