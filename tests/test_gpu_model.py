"""Gate 2b (GPU): the host-side mirror of fgpm()/predict()/simulate()/update() driving the CUDA path, against
(i) the values the reference prints in docs/reference/fgpm.html and (ii) the CPU oracle run with the SAME optimiser,
the same R-compatible RNG stream and the same start points ("same optimised hyper-parameters to optimiser tolerance")."""
import numpy as np
import pytest

from oracle import fungp_ref as ref
from oracle.rrng import RRng
from helpers import load_docs, docs_example_data, synth_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fungp_b200
    c = fungp_b200.Context(1)
    yield c
    c.close()


def test_docs_seeded_fits_on_gpu(ctx):
    """ms, mf, msf of the reference's first example, fitted one after another on one R RNG stream (set.seed(100))."""
    from fungp_b200 import model as M
    d = load_docs()
    sIn, fIn, sOut, rng = docs_example_data()
    ms = M.fgpm(sIn=sIn, sOut=sOut, rng=rng, ctx=ctx)
    mf = M.fgpm(fIn=fIn, sOut=sOut, rng=rng, ctx=ctx)
    msf = M.fgpm(sIn=sIn, fIn=fIn, sOut=sOut, rng=rng, ctx=ctx)
    assert abs(mf.negLogLik - d["mf"]["nll"]) < 5e-7
    np.testing.assert_allclose(mf.kern.f_lsHyps, d["mf"]["ls"], atol=6e-5)
    assert abs(mf.kern.varHyp - d["mf"]["var"]) < 6e-5
    assert abs(msf.negLogLik - d["msf"]["nll"]) < 5e-7
    np.testing.assert_allclose(np.r_[msf.kern.s_lsHyps, msf.kern.f_lsHyps], d["msf"]["ls"], atol=6e-5)
    assert abs(msf.kern.varHyp - d["msf"]["var"]) < 6e-5
    K = msf.preMats["L"] @ msf.preMats["L"].T                     # tcrossprod(m1@preMats$L), printed to 7 digits
    np.testing.assert_allclose(K, np.array(d["Ktt"]), rtol=0, atol=6e-7)
    assert ms.negLogLik < 28.0      # scipy's L-BFGS-B takes the neighbouring basin for `ms`
    for m in (ms, mf, msf):
        m.close()


def test_docs_custom_gauss_example_on_gpu(ctx):
    from fungp_b200 import model as M
    d = load_docs()
    sIn, fIn, sOut, rng = docs_example_data()
    m = M.fgpm(sIn=sIn, fIn=fIn, sOut=sOut, kerType="gauss", f_disType=["L2_byindex", "L2_bygroup"], f_pdims=[0, 5],
               f_basType=["NA", "B-splines"], rng=rng, ctx=ctx)
    assert abs(m.negLogLik - d["custom"]["nll"]) < 5e-7
    assert abs(m.kern.varHyp - d["custom"]["var"]) < 6e-5
    np.testing.assert_allclose(np.r_[m.kern.s_lsHyps, m.kern.f_lsHyps][:8], d["custom"]["ls_first8"], atol=6e-5)
    m.close()


@pytest.mark.parametrize("ker,n_starts", [("matern5_2", 1), ("gauss", 3), ("matern3_2", 2)])
def test_fit_predict_simulate_update_vs_oracle(ctx, ker, n_starts):
    from fungp_b200 import model as M
    n, m = 300, 40
    sIn, fIn, y, _ = synth_problem(77, n, 2, [30, 40], extra=m)
    tr, te = slice(0, n), slice(n, n + m)
    kw = dict(kerType=ker, f_disType=["L2_bygroup", "L2_byindex"], f_pdims=[4, 2], f_basType=["B-splines", "PCA"],
              nugget=1e-8, n_starts=n_starts, n_presample=20)
    g = M.fgpm(sIn=sIn[tr], fIn=[f[tr] for f in fIn], sOut=y[tr], rng=RRng(5), ctx=ctx, **kw)
    o = ref.fgpm(sIn=sIn[tr], fIn=[f[tr] for f in fIn], sOut=y[tr], rng=RRng(5), **kw)
    # same optimiser, same starts, nll identical to ~1e-12 => same trajectory up to FP noise.  L-BFGS-B stops on a
    # relative decrease of factr * eps = 2.2e-9, so the end point along a flat direction of the likelihood is only
    # determined to ~1e-4 relative while the optimum value itself agrees to 1e-8
    assert abs(g.negLogLik - o.negLogLik) <= 1e-8 * max(abs(o.negLogLik), n)
    np.testing.assert_allclose(np.r_[g.kern.s_lsHyps, g.kern.f_lsHyps], o.lsHyps, rtol=1e-4)
    assert abs(g.kern.varHyp - o.varHyp) <= 1e-4 * o.varHyp
    # predict / simulate with the ORACLE evaluated at the GPU model's hyper-parameters (isolates the algebra)
    o2 = ref.fgpm(sIn=sIn[tr], fIn=[f[tr] for f in fIn], sOut=y[tr], kerType=ker, f_disType=kw["f_disType"],
                  f_pdims=kw["f_pdims"], f_basType=kw["f_basType"], var_hyp=g.kern.varHyp, ls_s_hyp=g.kern.s_lsHyps,
                  ls_f_hyp=g.kern.f_lsHyps, nugget=1e-8)
    # PCA sign (quirk Q7) does not matter for distances: predictions agree
    pg = M.predict(g, sIn_pr=sIn[te], fIn_pr=[f[te] for f in fIn], detail="full")
    po = ref.predict(o2, sNew=sIn[te], fNew=[f[te] for f in fIn], detail="full")
    sd0 = np.sqrt(g.kern.varHyp)
    assert np.max(np.abs(pg["mean"] - po["mean"])) <= 1e-8 * max(1.0, np.max(np.abs(po["mean"])))
    assert np.max(np.abs(pg["sd"] - po["sd"])) <= 1e-8 * sd0 + 1e-8 * np.max(po["sd"])
    assert np.max(np.abs(pg["lower95"] - po["lower95"])) <= 1e-7 * max(1.0, np.max(np.abs(po["mean"])))
    assert np.max(np.abs(pg["K.pp"] - po["K.pp"])) <= 1e-12 * g.kern.varHyp * 10
    Z = RRng(1).rnorm(m * 6).reshape((6, m)).T
    sg = M.simulate(g, nsim=6, sIn_sm=sIn[te], fIn_sm=[f[te] for f in fIn], nugget_sm=1e-8, detail="full", Z=Z)
    so = ref.simulate(o2, Z, sNew=sIn[te], fNew=[f[te] for f in fIn], nugget_sm=1e-8, detail="full")
    assert np.max(np.abs(sg["sims"] - so["sims"])) <= 1e-6 * max(1.0, np.max(np.abs(so["sims"])))
    assert abs(M.loocv_q2(g) - ref.q2_loocv(o2)) <= 1e-8
    # update: substitute two outputs, keep hyper-parameters  (R/6_updating.R:91-323)
    y2 = y[tr].copy(); y2[[3, 10]] += 0.5
    gu = M.update(g, sOut_sb=y2[[3, 10]], ind_sb=[4, 11])
    ou = ref.fgpm(sIn=sIn[tr], fIn=[f[tr] for f in fIn], sOut=y2, kerType=ker, f_disType=kw["f_disType"], f_pdims=kw["f_pdims"],
                  f_basType=kw["f_basType"], var_hyp=g.kern.varHyp, ls_s_hyp=g.kern.s_lsHyps, ls_f_hyp=g.kern.f_lsHyps, nugget=1e-8)
    assert np.max(np.abs(gu.preMats["LInvY"] - ou.LInvY)) <= 1e-8 * np.max(np.abs(ou.LInvY))
    # update: delete 5 points and re-estimate the variance only
    gd = M.update(g, ind_dl=[1, 2, 3, 4, 5], var_re=True)
    keep = slice(5, n)
    od = ref.fgpm(sIn=sIn[keep], fIn=[f[keep] for f in fIn], sOut=y[keep], kerType=ker, f_disType=kw["f_disType"],
                  f_pdims=kw["f_pdims"], f_basType=kw["f_basType"], ls_s_hyp=g.kern.s_lsHyps, ls_f_hyp=g.kern.f_lsHyps, nugget=1e-8)
    assert abs(gd.kern.varHyp - od.varHyp) <= 1e-7 * od.varHyp
    for mm in (g, gu, gd):
        mm.close()
