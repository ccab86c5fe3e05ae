"""C++ host mirror on a box WITHOUT a GPU: argument handling works, and anything that needs a
table fails loudly (exit code 3, message on stderr) -- there is no CPU fallback behind `bn`/`mn`."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bnpp_b200", "bin")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "bn")), reason="host CLIs not built (run build())")


def test_usage_and_bad_flag():
    p = subprocess.run([os.path.join(BIN, "bn")], capture_output=True, text=True)
    assert p.returncode == 1 and p.stdout.startswith("usage: ")
    p = subprocess.run([os.path.join(BIN, "bn"), "x.uai", "-h"], capture_output=True, text=True)
    assert p.returncode == 0 and "-wmf\tvariable elimination using weighted min-fill heuristic" in p.stdout
    p = subprocess.run([os.path.join(BIN, "bn"), "x.uai", "-zz"], capture_output=True, text=True)
    assert p.returncode == 255 and "Error: invalid option `-zz'." in p.stderr      # exit(-1), code/bn.cpp:201-205
    p = subprocess.run([os.path.join(BIN, "bn"), "/nonexistent/model.uai", "-pr"], capture_output=True, text=True)
    assert p.returncode == 255 and "Error: couldn't read file /nonexistent/model.uai" in p.stderr


def test_wrong_network_kind(tmp_path, golden_models):
    bayes = tmp_path / "asia.uai"
    bayes.write_text(golden_models["asia"]["uai"])
    p = subprocess.run([os.path.join(BIN, "mn"), str(bayes), str(bayes)], capture_output=True, text=True, input="quit\n")
    assert p.returncode == 255 and "is not a MARKOV net." in p.stderr               # code/io.cpp:139-143, returns -2 -> main -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="this box has a GPU")
def test_no_cpu_fallback(tmp_path, golden_models):
    f = tmp_path / "asia.uai"
    f.write_text(golden_models["asia"]["uai"])
    p = subprocess.run([os.path.join(BIN, "bn"), str(f), "-pr"], capture_output=True, text=True)
    assert p.returncode == 3
    assert "cannot create a CUDA context" in p.stderr and "Partition" not in p.stdout


@pytest.mark.skipif(torch.cuda.is_available(), reason="this box has a GPU")
def test_api_harness_links_and_has_no_cpu_fallback(tmp_path, golden_models):
    """tests/bin/harness = the reference's own harness source compiled against the product's headers
    (tests/harness/Makefile): it must find libbnpp_b200.so next to the package and refuse to compute without a GPU"""
    exe = os.path.join(ROOT, "tests", "bin", "harness")
    if not os.path.exists(exe):
        pytest.skip("harness not built (run build())")
    f = tmp_path / "asia.uai"
    f.write_text(golden_models["asia"]["uai"])
    p = subprocess.run([exe], input="model %s\npr\n" % f, capture_output=True, text=True)
    assert p.returncode == 3, (p.returncode, p.stderr[-300:])
    assert "cannot create a CUDA context" in p.stderr and "PR" not in p.stdout
