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


def _check(paths):
    exe = os.path.join(ROOT, "tests", "bin", "uai_parse_check")
    if not os.path.exists(exe):
        pytest.skip("uai_parse_check not built (run build())")
    p = subprocess.run([exe] + [str(x) for x in paths], capture_output=True, text=True)
    rows = {}
    for line in p.stdout.splitlines():
        f = line.split()
        rows[os.path.basename(f[0])] = f[5]
    return p.returncode, rows


def test_fast_uai_reader_equals_reference_reader(tmp_path, golden_models):
    """bnpp_b200/host/uai_parse.hpp (one read, tokens in place, from_chars) against a restatement of the reference's
    reader (code/io.cpp:14-100): bit-identical values and partitions on every shipped network and on the
    golden texts; anything unusual must make the fast reader step aside (FALLBACK), never differ"""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "models", "*", "*.uai")))
    for name, m in golden_models.items():
        f = tmp_path / (name + ".golden.uai")
        f.write_text(m["uai"])
        files.append(str(f))
    rc, rows = _check(files)
    assert rc == 0
    assert len(rows) >= len(golden_models) and all(v == "SAME" for v in rows.values()), rows
    body = "2\n2 3\n2\n1 0\n2 0 1\n2 0.25 0.75\n6 1 2 3 4 5 6\n"
    odd = {
        "comments.uai": "BAYES # a comment\n# whole line\n2 #cards follow\n2 3\n2\n1 0\n2 0 1\n2 0.25 0.75 # tail\n6 1 2 3 4 5 6 #end",
        "crlf.uai": ("MARKOV\n" + body).replace("\n", "\r\n"),
        "forms.uai": "BAYES\n" + body.replace("0.25 0.75", ".25 7.5e-1").replace("1 2 3", "1. 2E0 -0"),
        "plus.uai": "BAYES\n" + body.replace("0.25", "+0.25"),
        "subnormal.uai": "BAYES\n" + body.replace("0.25", "1e-320"),
        "huge.uai": "BAYES\n" + body.replace("0.25", "1e400"),
        "hexfloat.uai": "BAYES\n" + body.replace("0.25", "0x1p-2"),
        "junk_int.uai": "BAYES\n" + body.replace("2 3", "2 3abc", 1),
        "truncated.uai": "BAYES\n" + body[:-8],
        "bad_scope.uai": "BAYES\n" + body.replace("2 0 1", "2 0 7"),
        "header.uai": "BAYESIAN\n" + body,
        "empty.uai": "",
    }
    paths = []
    for name, text in odd.items():
        f = tmp_path / name
        f.write_bytes(text.encode())
        paths.append(f)
    rc, rows = _check(paths)
    assert rc == 0, rows
    assert rows["comments.uai"] == "SAME" and rows["crlf.uai"] == "SAME" and rows["forms.uai"] == "SAME", rows
    for name in ("plus.uai", "subnormal.uai", "huge.uai", "hexfloat.uai", "junk_int.uai", "truncated.uai", "bad_scope.uai",
                 "header.uai", "empty.uai"):
        assert rows[name] == "FALLBACK", (name, rows[name])


def test_uai_solution_writers_cpu(tmp_path):
    """SURVEY 8f row 3 without a device: write_uai_pr gives the shipped grid3x3.uai.PR byte for byte (PR / 1 / 14.8899,
    6 significant digits of log10 Z), write_uai_evidence round-trips through read_uai_evidence (code/io.cpp:157-180)"""
    exe = os.path.join(ROOT, "tests", "bin", "uai_write_check")
    if not os.path.exists(exe):
        pytest.skip("uai_write_check not built (run build())")
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert out.stdout == "PR\n1\n14.8899\n--\n1\n3 0 1 4 1 5 1\n--\nroundtrip ok\n", out.stdout
    assert open(str(tmp_path / "w.PR")).read().split() == ["PR", "1", "14.8899"]
