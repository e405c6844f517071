"""The parameter file text used by tests/test_zz_gpu_high_degree.py IS the reference's shipped
parameters.prm (comments aside) - checked against the real file where /root/reference exists."""
import os

import pytest

REF = "/root/reference/parameters.prm"


def effective_lines(text):
    return [" ".join(l.split()) for l in text.splitlines() if l.strip() and not l.strip().startswith("#")]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_shipped_prm_string_equals_the_reference_file():
    from test_zz_gpu_high_degree import SHIPPED_PARAMETERS_PRM
    assert effective_lines(SHIPPED_PARAMETERS_PRM) == effective_lines(open(REF).read())


def test_cpp_parser_accepts_the_shipped_file(tmp_path):
    """host/parameters.cc parses it: the run then stops at the device set-up on a machine without
    a GPU (exit code 1 from gf_create), with one on the missing participant configuration - in
    both cases AFTER the parameter summary, never with a parser complaint."""
    import subprocess
    from dealii_adapter_b200 import build
    from test_zz_gpu_high_degree import SHIPPED_PARAMETERS_PRM
    exe = build.build_elasticity()[0]
    (tmp_path / "parameters.prm").write_text(SHIPPED_PARAMETERS_PRM)
    r = subprocess.run([exe, "parameters.prm"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 1
    assert "no such" not in r.stderr and "Polynomial degree" not in r.stderr
