"""The reference's own transi test program (tests/transi/transi_test_program.c) compiled UNCHANGED against
include/ectrans/transi.h and linked with the CUDA library.

CPU part (runs where /root/reference exists): the program and a generated member-for-member check of struct Trans_t
compile with gcc.  GPU part: the prebuilt binary (oracle/_ref/, travels with the snapshot) runs and reproduces the known
answers of the reference test: a constant grid-point field c has psi(0,0) = c as its only coefficient
(transi_test_program.c:76-81,150-164) and comes back as c."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref", "transi_test_program")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "tests/transi/transi_test_program.c")), reason="reference tree not present")
def test_reference_transi_program_compiles(built):
    subprocess.check_call(["make", "-C", ROOT, "transi_ref"])
    assert os.path.exists(BIN)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src/transi/transi.h")), reason="reference tree not present")
def test_trans_t_has_every_reference_member(built, tmp_path):
    """Every member of the reference's struct Trans_t and every function it declares exists with a compatible type:
    a C file that touches all of them (generated from the reference header) compiles against ours."""
    src = open(os.path.join(REF, "src/transi/transi.h")).read()
    body = src[src.index("struct Trans_t {"):]
    body = body[:body.index("\n};")]
    members = re.findall(r"^\s*(?:const\s+)?(?:int|double|_bool|char|void|size_t)\s*\*?\s*(\w+)\s*;", body, re.M)
    assert len(members) == 67, len(members)            # transi.h:701-850
    funcs = re.findall(r"^(?:int|const char\*)\s+(trans_\w+)\s*\(", src, re.M) + re.findall(r"^struct \w+ (new_\w+)\s*\(", src, re.M)
    assert len(funcs) >= 40, len(funcs)
    c = ['#include "ectrans/transi.h"', "int main(void) {", "  struct Trans_t t; (void)sizeof(t);"]
    c += [f"  (void)sizeof(t.{m});" for m in members]
    c += [f"  (void)&{f};" for f in funcs]
    c += ["  return 0;", "}"]
    f = tmp_path / "members.c"
    f.write_text("\n".join(c))
    lib = os.path.join(ROOT, "ectrans_b200", "lib")
    subprocess.check_call(["gcc", "-std=gnu99", "-I", os.path.join(ROOT, "include"), str(f), "-o", str(tmp_path / "members"),
                           "-L", lib, "-lectrans_b200", f"-Wl,-rpath,{lib}"])


@pytest.mark.gpu
def test_reference_transi_program_runs():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/transi_test_program not built (needs the reference tree at build time)")
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=600, env=dict(os.environ, TRANS_USE_MPI="0"))
    assert out.returncode == 0, out.stderr[-3000:]
    err = out.stderr
    # distgrid: no deviation lines beyond the first three samples of each field
    sc = re.findall(r"rspscalar\[(\d+)\]\[(\d+)\] : ([-\d.]+)", err)
    big = {(int(j), int(i)): float(v) for j, i, v in sc if abs(float(v)) > 1e-5}
    assert big == {(0, 0): 3.0, (1, 0): 4.0}, big                 # constant fields 3 and 4: only psi(0,0)
    g = {(int(j), int(i)): float(v) for j, i, v in re.findall(r"rgpg\[(\d+)\]\[(\d+)\] : ([-\d.]+)", err)}
    assert abs(g[(0, 0)] - 1.0) < 1e-5 and abs(g[(1, 0)] - 2.0) < 1e-5 and abs(g[(2, 0)] - 3.0) < 1e-5
    assert "nprtrw = 1" in err
