"""The R .Call glue cannot be linked here (no R in the image): type-check it against the mock declarations and make
sure it binds only symbols that include/fungp_b200.h declares and registers every wrapper it defines."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "rglue")


def test_glue_type_checks_against_mock_headers():
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-I" + os.path.join(G, "mock"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(G, "fungp_glue.c"), os.path.join(G, "init.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "warning" not in r.stderr, r.stderr


def test_glue_uses_only_declared_c_abi_and_registers_all_wrappers():
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fungp_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(fgp_[a-z0-9_]+)\s*\(", h))
    glue = open(os.path.join(G, "fungp_glue.c")).read()
    glue_nc = re.sub(r"/\*.*?\*/", "", glue, flags=re.S)
    used = set(re.findall(r"\b(fgp_[a-z0-9_]+)\s*\(", glue_nc))
    assert used <= declared, used - declared
    wrappers = set(re.findall(r"^SEXP (fgpR_[a-z0-9_]+)\(", glue_nc, flags=re.M))
    init = open(os.path.join(G, "init.c")).read()
    registered = set(re.findall(r'\{"(fgpR_[a-z0-9_]+)"', init))
    assert wrappers == registered and len(wrappers) >= 13
    rfile = open(os.path.join(G, "R", "zz_b200_backend.R")).read()
    called = set(re.findall(r"\.Call\((fgpR_[a-z0-9_]+)", rfile))
    assert called <= wrappers, called - wrappers
