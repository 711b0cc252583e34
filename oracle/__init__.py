"""CPU oracle for the funGp hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, op for op in numpy/scipy float64, the algorithm of the
reference R package funGp 1.0.0 for the path named in BASELINE.json
(dimReduction -> distance tensors -> correlation assembly -> concentrated
negative log-likelihood -> prediction / simulation solves -> LOOCV fitness).

It is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``fungp_b200/`` imports it, and
the product path raises when the CUDA library is missing.

Parity pinning: the restatement is checked (tests/test_oracle_golden.py)
against every known-answer artefact the reference ships for this path --
the five stored models in data/precalculated_Xfgpm_objects.rda (full double
precision nll, sigma^2, L, LInvY, Q2) and the rendered example output in
docs/reference/fgpm.html.  predict()/simulate() numerics are NOT pinned by any
reference artefact ("parity unpinned" for those two: CUDA is compared with this
restatement only).  R itself is not installed in the build image, so the real
reference cannot be executed; see DESIGN.md.
"""
