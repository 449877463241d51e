"""GPU: the reference's own RuntimeVisitor (compiled unchanged from /root/reference/src into oracle/_ref/libabc_ref.a)
driving the C++ drop-in CudaCiphertextFactory (abc_b200/cpp) — the reference's SEAL-backed test cases with the
factory swapped, and the batched end-to-end programs of SURVEY.md 8(d)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "abc_b200", "bin", "abc_driver")


def run(*args):
    assert os.path.exists(DRIVER), "abc_driver is not built (python -c 'import __graft_entry__ as g; g.build()')"
    p = subprocess.run([DRIVER, *args], capture_output=True, text=True, timeout=600)
    print(p.stdout[-4000:], p.stderr[-2000:])
    summary = json.loads(p.stdout.strip().splitlines()[-1])
    assert p.returncode == 0 and summary["failures"] == 0, p.stdout[-4000:]
    return summary, p.stdout


def test_reference_runtime_visitor_and_factory_kats():
    summary, _ = run("kats")
    assert summary["cases"] >= 25


def test_batched_programs_n8192():
    summary, out = run("programs", "8192")
    assert summary["cases"] == 4
    for name in ("HammingDistance", "L2Distance", "BoxBlur", "GxKernel"):
        assert "[ ok ] program." + name in out


def test_batched_programs_n16384():
    # 8192 slots per row is not a square image: distance programs only
    summary, _ = run("programs", "16384")
    assert summary["cases"] == 2


def test_lock_step_batch_driver():
    """ONE RuntimeVisitor walk drives B instances (the per-instance `secret` inputs come from the tables
    registered with the factory); every instance of L2Distance and sampled instances of the stencils are checked against
    the plain evaluation inside the driver."""
    summary, out = run("programs", "8192", "--batch", "24", "--steps", "2")
    assert summary["cases"] == 3
    for name in ("L2Distance", "BoxBlur", "GxKernel"):
        assert "[ ok ] batch." + name in out
    assert '"all_instances_checked": true' in out


def test_a16_and_unsupported_ops_are_in_the_kats():
    _, out = run("kats")
    assert "[ ok ] a16.Cleartext<int>::subtract_inplace(ciphertext)" in out
    assert out.count("[ ok ] unsupported.") >= 15


def test_transparent_results_throw_like_seal_when_asked_to():
    """std::logic_error("result ciphertext is transparent") from the C++ drop-in (setThrowOnTransparent),
    through factory ops and through the reference's RuntimeVisitor; off by default."""
    _, out = run("kats")
    assert out.count("[ ok ] transparent.") == 4


def test_demo_csv_breakdown(tmp_path):
    """The reference demo's CSV schema (examples/main.cpp:41), measured: one row of four phase times in ms."""
    target = tmp_path / "demo.csv"
    summary, _ = run("demo", str(target), "8192", "--batch", "8")
    assert summary["cases"] == 1
    header, row = target.read_text().strip().splitlines()
    assert header == "t_keygen,t_input_encryption,t_computation,t_decryption"
    vals = [float(v) for v in row.split(",")]
    assert len(vals) == 4 and all(v > 0 for v in vals)


def test_caller_supplied_sampler_key_shares_keys_between_factories():
    _, out = run("kats")
    assert out.count("[ ok ] rngkey.") == 2
