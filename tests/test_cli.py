"""Boundary B0: the reference's Makefile flow (`make` -> `pcal2tla *tla` ; `tlc *tla`, Makefile:1-7) driven
by this repo's bin/ wrappers, on builder-authored PlusCal specs of the same shape (models/demo)."""
import os
import shutil
import subprocess
import sys
import tempfile

import pytest

from conftest import ROOT


def _run_make(target="all"):
    d = tempfile.mkdtemp(prefix="tlag_cli_")
    for f in os.listdir(os.path.join(ROOT, "models", "demo")):
        shutil.copy(os.path.join(ROOT, "models", "demo", f), d)
    env = dict(os.environ)
    env["PATH"] = os.path.join(ROOT, "bin") + os.pathsep + os.path.dirname(sys.executable) + os.pathsep + env["PATH"]
    p = subprocess.run(["make", target], cwd=d, env=env, capture_output=True, text=True, timeout=300)
    return d, p


def test_pcal2tla_in_place_translation_cfg_and_backup():
    d, p = _run_make("transpile")
    assert p.returncode == 0, p.stderr
    txt = open(os.path.join(d, "race.tla")).read()
    assert "\\* BEGIN TRANSLATION" in txt and "\\* END TRANSLATION" in txt and "Spec == Init /\\ [][Next]_vars" in txt
    assert os.path.exists(os.path.join(d, "race.old")) and os.path.exists(os.path.join(d, "lock.old"))
    # idempotent: translating again replaces the block instead of appending a second one
    _ = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "pcal2tla"), os.path.join(d, "race.tla")], check=True,
                       capture_output=True)
    assert open(os.path.join(d, "race.tla")).read().count("BEGIN TRANSLATION") == 1


def test_tlc_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d, p = _run_make("all")
    assert p.returncode != 0                       # make stops: no CPU fallback exists
    assert "CUDA" in (p.stdout + p.stderr)


@pytest.mark.gpu
def test_make_flow_on_gpu_reports_like_tlc():
    d, p = _run_make("all")
    out = p.stdout
    assert p.returncode != 0                       # race.tla violates its invariant -> make stops (Makefile:6-7)
    # lock.tla (checked first: glob order) passes
    assert "Model checking completed. No error has been found." in out
    assert "45 states generated, 26 distinct states found, 0 states left on queue." in out
    assert "The depth of the complete state graph search is 8." in out
    # race.tla: TLC-format counterexample (README.md:267-321 layout)
    assert "Error: Invariant Correct is violated." in out
    assert "State 1: <Initial predicate>" in out and "/\\ counter = 0" in out
    assert "State 5: <Action line" in out and "of module race>" in out
    assert '/\\ pc = <<"Done", "Done">>' in out and "/\\ counter = 1" in out
