"""Host-side checkers used by the GPU benches, pinned on the CPU: the stencil expectation of tools/stencil10k_bench.py
against a restatement of the reference's naiveBoxBlur / naiveGxKernel (test/end-to-end/BoxBlurTest.cpp:21-38,
GxKernelTest.cpp:20-37), and its NAF key-switch count against the oracle's rotation plan."""
import numpy as np

from tools import stencil10k_bench as sb


def naive_stencil(img, w):
    """The reference's loops, including the index arithmetic: ((x+i)*imgSize + (y+j)) is an int that converts to size_t
    before `% img.size()` (2^64 = 0 mod 4096, so negative indices wrap cyclically)."""
    size = int(np.ceil(np.sqrt(len(img))))
    out = list(img)
    for x in range(size):
        for y in range(size):
            value = 0
            for j in (-1, 0, 1):
                for i in (-1, 0, 1):
                    idx = (((x + i) * size + (y + j)) % (1 << 64)) % len(img)
                    value += w[i + 1][j + 1] * img[idx]
            out[size * x + y] = value
    return out


def test_stencil_expectation_is_the_references_naive_functions():
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 1025, size=(2, sb.SIZE * sb.SIZE), dtype=np.int64)
    for w in (sb.BOX, sb.GX):
        got = sb.expected(imgs, w)
        for b in range(2):
            assert list(got[b]) == naive_stencil([int(v) for v in imgs[b]], w)


def test_naf_weight_is_the_oracles_key_switch_count(oracle4096):
    o = oracle4096
    # (steps whose NAF has a digit N/2 are left out: that digit is the identity on a row and SEAL skips it)
    for step in list(range(-130, 131)) + [-1023, 1023, 511, -767, 341, -683]:
        assert sb.naf_weight(step) == o.rotate_keyswitch_count(step), step
    assert sb.key_switches(sb.BOX) == 12 and sb.key_switches(sb.GX) == 10
    assert [k for k, _ in sb.taps(sb.BOX)] == [-65, -64, -63, -1, 0, 1, 63, 64, 65]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: the reference's RuntimeVisitor over the CPU port, one JSON line with the
    keys of the bench contract (impl, metric, unit, value, config identical to our arm's, cpu_baseline, e2e with zero copy
    bytes).  Tiny sample: one instance per host process, one step."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-instances-per-core", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["higher_is_better"] is True and r["unit"] == "ops/s" and r["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in r
    assert r["config"]["workload"].startswith("l2distance_batched") and "batch_per_gpu" in r["config"]
    assert r["cpu_baseline"]["kind"] in ("port", "reference") and r["cpu_baseline"]["cores"] >= 1
    assert r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_host_number_theory_matches_the_oracle(tmp_path):
    """abc_b200/csrc/hostmath.hpp is what abc_ctx_create builds its tables from (SEAL's default primes, the plain modulus,
    minimal primitive 2N-th roots, the 61-bit auxiliary primes, inverses, Barrett ratios).  Compiled alone with g++ (no
    GPU) and compared with the oracle, N = 4096 ... 32768."""
    import os
    import subprocess
    from oracle.bfv_oracle import Oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "hostmath_probe")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(root, "tests", "host", "hostmath_probe.cpp")], check=True)
    for N in (4096, 8192, 16384, 32768):
        out = subprocess.run([exe, str(N)], capture_output=True, text=True, check=True).stdout.split("\n")
        vals = {}
        for line in out:
            if line:
                k, *v = line.split()
                vals.setdefault(k, []).append([int(x) for x in v])
        o = Oracle(N)
        assert vals["k"][0][0] == o.k
        assert [q[0] for q in vals["q"]] == o.primes
        assert vals["t"][0][0] == o.t
        assert [p[0] for p in vals["psi"]] == [o.psi(i) for i in range(o.k)]
        msk, gamma, B = o.aux_primes()
        aux = [a[0] for a in vals["aux"]]
        assert set(B) | {msk} <= set(aux) or aux[:len(B)] == B        # SEAL takes its base from this prime sequence
        q0, q1 = o.primes[0], o.primes[1]
        assert vals["inv"][0][0] == pow(q0 % q1, -1, q1)
        hi, lo = vals["barrett"][0]
        assert (hi << 64 | lo) == (1 << 128) // q0
        assert vals["brev"][0][0] == int("{:013b}".format(0x1234)[::-1], 2)
