"""bench.py host logic that needs no GPU: the byte models of DESIGN.md 3.3 / SURVEY.md 8(d) and the
reference arm (`--impl reference`: the CPU oracle port on the host cores) with its JSON contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_byte_models_are_consistent():
    n, e, d = 1000, 7000, 64
    for conv in ("gt", "gat"):
        k = bench.kernel_alg_bytes(conv, n, e, d)
        c = bench.kernel_compulsory_bytes(conv, n, n, e, d)
        assert set(k) == set(c) == {"fwd", "bwd_row", "bwd_col"}
        # gathers re-read neighbour rows: the gather model is never below the compulsory bytes on a graph
        # with more edges than nodes
        assert all(k[p] >= c[p] * 0.99 for p in k)
        assert bench.alg_bytes(conv, "fwd+bwd", n, e, d) == bench.alg_bytes(conv, "fwd", n, e, d) + bench.alg_bytes(conv, "bwd", n, e, d)
    # GT forward 8Ed + 8Nd + 4E + 4(N+1)
    assert bench.alg_bytes("gt", "fwd", n, e, d) == 8.0 * e * d + 8.0 * n * d + 4.0 * e + 4.0 * (n + 1)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cora-gt",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["value"] > 0 and line["higher_is_better"] is True and line["config"]["workload"] == "cora-gt"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_tf32_split_properties():
    """The 3xTF32 split of csrc/tc_common.cuh (round_tf32 / split4), restated in numpy: hi has 10 mantissa
    bits, hi + lo == x exactly in fp32, |lo| <= 2^-11 |x| (rounding, not truncation), and the three-term
    product a_hi b_hi + a_hi b_lo + a_lo b_hi is within 2^-21 |a b| of the exact product."""
    import numpy as np
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    bits = x.view(np.uint32)
    hi = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = x - hi
    assert np.all((hi.view(np.uint32) & np.uint32(0x1FFF)) == 0)
    assert np.array_equal(hi + lo, x)
    assert np.all(np.abs(lo) <= np.abs(x) * 2.0 ** -11 * (1 + 1e-6))
    a, b = x[:100000].astype(np.float64), x[100000:].astype(np.float64)
    ah, al = hi[:100000].astype(np.float64), lo[:100000].astype(np.float64)
    bh, bl = hi[100000:].astype(np.float64), lo[100000:].astype(np.float64)
    approx = ah * bh + ah * bl + al * bh
    assert np.all(np.abs(approx - a * b) <= np.abs(a * b) * 2.0 ** -21)
