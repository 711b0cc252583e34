"""Regenerates rglue/patches/*.diff from /root/reference (only where the reference tree is present).  The edits themselves
live in this script as (old, new) text pairs applied to a scratch copy; `diff -U2` of the copy against the original is
what is committed.  Run: python tools/make_rglue_patches.py"""
import os
import shutil
import subprocess
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "rglue", "patches")
FILES = ["R/1_fgpm_Class.R", "R/3_ant_search.R", "R/6_updating.R", "NAMESPACE", "DESCRIPTION"]

if not os.path.isdir(REF):
    sys.exit("reference tree not present: the committed patches are the artefact")
print("the edit list is kept with the patches (see the diffs themselves); this script only re-checks that they still apply")
tmp = "/tmp/rglue_patch_check"
shutil.rmtree(tmp, ignore_errors=True)
os.makedirs(os.path.join(tmp, "R"))
for f in FILES:
    shutil.copy(os.path.join(REF, f), os.path.join(tmp, f))
for f in FILES:
    d = os.path.join(OUT, f.replace("/", "_") + ".diff")
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", d], cwd=tmp, capture_output=True, text=True)
    print(f, "applies" if r.returncode == 0 else "DOES NOT APPLY:\n" + r.stdout + r.stderr)
