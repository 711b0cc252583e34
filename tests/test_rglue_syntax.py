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


def test_patches_apply_to_the_reference_tree(tmp_path):
    """rglue/patches/*.diff wire the *_B200 helpers into fgpm(), predict.fgpm, simulate.fgpm, update and run_ACO: they must
    apply cleanly to the reference sources (only checkable where /root/reference exists), and every helper they call must
    be defined in zz_b200_backend.R."""
    import glob
    import shutil
    patches = sorted(glob.glob(os.path.join(G, "patches", "*.diff")))
    assert len(patches) >= 5
    rfile = open(os.path.join(G, "R", "zz_b200_backend.R")).read()
    defined = set(re.findall(r"^([A-Za-z_.][A-Za-z0-9_.]*) <- function", rfile, flags=re.M))
    for p in patches:
        added = "\n".join(l[1:] for l in open(p).read().splitlines() if l.startswith("+") and not l.startswith("+++"))
        for name in set(re.findall(r"\b([A-Za-z_.]+_B200|b200_[a-z_]+)\(", added)):
            assert name in defined, (os.path.basename(p), name)
    ref = "/root/reference"
    if not os.path.isdir(ref):
        return
    for f in ("R/1_fgpm_Class.R", "R/3_ant_search.R", "R/6_updating.R", "NAMESPACE", "DESCRIPTION"):
        os.makedirs(os.path.dirname(os.path.join(tmp_path, f)), exist_ok=True)
        shutil.copy(os.path.join(ref, f), os.path.join(tmp_path, f))
    for p in patches:
        r = subprocess.run(["patch", "-p1", "--dry-run", "-i", p], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, (p, r.stdout, r.stderr)
