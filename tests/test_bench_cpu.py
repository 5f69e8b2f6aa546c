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
