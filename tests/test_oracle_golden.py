"""Gate 1 (CPU): the oracle restatement against every known-answer artefact of the reference."""
import numpy as np
import pytest

from oracle import fungp_ref as ref
from oracle.rrng import RRng
from helpers import load_models, load_docs, docs_example_data, MODEL_NAMES


def test_r_rng_matches_printed_runif():
    d = load_docs()
    np.testing.assert_allclose(RRng(100).runif(5), d["runif_first5"], atol=5e-9)
    # set.seed(1); rnorm(3) -- base R documentation value
    np.testing.assert_allclose(RRng(1).rnorm(3), [-0.6264538, 0.1836433, -0.8356286], atol=5e-8)


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_stored_models_full_precision(name):
    """data/precalculated_Xfgpm_objects.rda: coefs, nll, sigma^2, L, LInvY, Q2 at stored hypers."""
    g = load_models()[name]
    fIn = g.get("fIn")
    bcj = ref.dimReduction(fIn, g["f_pdims"], g["f_basType"]) if fIn else None
    Ms = []
    if g["ds"]:
        Ms += ref.setDistMatrix_S(g["sIn"], g["sIn"])
    if fIn:
        for c, cg, B, Bg in zip(bcj["coefs"], g["coefs"], bcj["basis"], g["basis"]):
            np.testing.assert_allclose(B, Bg, atol=1e-14)
            np.testing.assert_allclose(c, cg, atol=1e-13)
        Ms += ref.setDistMatrix_F(bcj["coefs"], bcj["coefs"], bcj["J"], g["f_disType"])
    theta = np.concatenate([g["s_ls"], g["f_ls"]])
    nll, sig2 = ref.negLogLik(theta, [], Ms, g["sOut"], g["kerType"], g["nugget"])
    assert abs(nll - g["negLogLik"]) <= 1e-12 * abs(g["negLogLik"])
    assert abs(sig2 - g["varHyp"]) <= 1e-11 * g["varHyp"]
    L, LInvY = ref.preMats(Ms, g["sOut"], g["varHyp"], theta, g["kerType"], g["nugget"])
    assert np.max(np.abs(L - g["L"])) <= 1e-11 * np.max(np.abs(g["L"]))
    assert np.max(np.abs(LInvY - g["LInvY"])) <= 1e-10 * np.max(np.abs(g["LInvY"]))
    q2 = ref.q2(g["sOut"], ref.getOut_loocv(g["L"], g["varHyp"], g["nugget"], g["sOut"]))
    assert abs(q2 - g["fitness"]) <= 1e-12


def test_docs_bspline_basis_and_coefs():
    d = load_docs()
    sIn, fIn, sOut, _ = docs_example_data()
    bcj = ref.dimReduction(fIn, [3, 3], ["B-splines", "B-splines"])
    for i in range(2):
        np.testing.assert_allclose(bcj["basis"][i], np.array(d["basis"][i]), rtol=0, atol=6e-8)
        np.testing.assert_allclose(bcj["coefs"][i], np.array(d["coefs"][i]), rtol=0, atol=6e-8)


def _docs_model(kind):
    sIn, fIn, sOut, rng = docs_example_data()
    bcj = ref.dimReduction(fIn, [3, 3], ["B-splines", "B-splines"])
    sMs = ref.setDistMatrix_S(sIn, sIn)
    fMs = ref.setDistMatrix_F(bcj["coefs"], bcj["coefs"], bcj["J"], ["L2_bygroup"] * 2)
    return sMs, fMs, sOut


def test_docs_hybrid_optimum_is_on_bounds_and_Ktt():
    """msf: nll 2.841058 at theta = upper bounds (Q8); tcrossprod(L) printed to 7 digits."""
    d = load_docs()
    sMs, fMs, y = _docs_model("msf")
    Ms = sMs + fMs
    ub = ref.setBounds(Ms)[1]
    np.testing.assert_allclose(ub, d["msf"]["ls"], atol=5e-5)
    nll, sig2 = ref.negLogLik(ub, [], Ms, y, "matern5_2", 1e-8)
    assert abs(nll - d["msf"]["nll"]) < 5e-7
    assert abs(sig2 - d["msf"]["var"]) < 5e-5
    L, _ = ref.preMats(Ms, y, sig2, ub, "matern5_2", 1e-8)
    K = L @ L.T
    np.testing.assert_allclose(K, np.array(d["Ktt"]), rtol=0, atol=6e-7)


def test_docs_seeded_fits_end_to_end():
    """Whole fgpm() flow incl. RNG stream: ms, mf, msf fitted one after another on one stream."""
    d = load_docs()
    sIn, fIn, sOut, rng = docs_example_data()
    ms = ref.fgpm(sIn=sIn, sOut=sOut, rng=rng)      # consumes runif(2*20); different basin under scipy
    mf = ref.fgpm(fIn=fIn, sOut=sOut, rng=rng)
    msf = ref.fgpm(sIn=sIn, fIn=fIn, sOut=sOut, rng=rng)
    assert abs(mf.negLogLik - d["mf"]["nll"]) < 5e-7
    np.testing.assert_allclose(mf.lsHyps, d["mf"]["ls"], atol=6e-5)
    assert abs(mf.varHyp - d["mf"]["var"]) < 6e-5
    assert abs(msf.negLogLik - d["msf"]["nll"]) < 5e-7
    np.testing.assert_allclose(msf.lsHyps, d["msf"]["ls"], atol=6e-5)
    assert abs(msf.varHyp - d["msf"]["var"]) < 6e-5
    assert ms.negLogLik < 28.0  # scipy's L-BFGS-B lands in the neighbouring basin (27.39) -- documented


def test_docs_custom_gauss_example():
    """gauss + L2_byindex (no projection, 10 ls) + L2_bygroup B-splines p=5: nll 18.777358."""
    d = load_docs()
    sIn, fIn, sOut, rng = docs_example_data()
    m = ref.fgpm(sIn=sIn, fIn=fIn, sOut=sOut, kerType="gauss", f_disType=["L2_byindex", "L2_bygroup"],
                 f_pdims=[0, 5], f_basType=["NA", "B-splines"], rng=rng)
    assert abs(m.negLogLik - d["custom"]["nll"]) < 5e-7
    assert abs(m.varHyp - d["custom"]["var"]) < 6e-5
    np.testing.assert_allclose(m.lsHyps[:8], d["custom"]["ls_first8"], atol=6e-5)
