"""The reference arm of bench.py (`--impl reference`) needs no GPU: it times the reference's own ikd-Tree under the
restated update loop on the host cores.  Its JSON line is what the driver parses, so the keys are pinned here."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_line_has_the_contract_keys():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--map-points", "200000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "scans/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["metric"].startswith("scans/sec for IESKF update")
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


def test_knn_stream_model_counts_the_candidates_of_a_3x3x3_block():
    """bench.py's candidate-streaming model (B_knn_stream = 16 + 16 C + 20) against a brute-force count."""
    import numpy as np

    sys.path.insert(0, str(ROOT))
    import bench

    rng = np.random.default_rng(7)
    mp = rng.uniform(-20, 20, (40000, 3)).astype(np.float32)
    q = rng.uniform(-10, 10, (500, 3)).astype(np.float32)
    grid = bench.cell_count_grid(mp, cell=1.5)
    m = bench.knn_stream_model(grid, q, sample=len(q))
    inv = np.float32(1.0 / 1.5)
    ck = np.floor(mp * inv).astype(np.int64)
    qc = np.floor(q * inv).astype(np.int64)
    cand = np.array([np.count_nonzero(np.all(np.abs(ck - c) <= 1, axis=1)) for c in qc])
    assert m["queries_sampled"] == len(q)
    assert abs(m["candidates_per_query_mean"] - cand.mean()) < 1e-9
    assert abs(m["B_knn_stream_bytes_per_query"] - (36 + 16 * cand.mean())) < 1e-6
    assert m["B_bursts_bytes_per_query"] >= 27 * 64
    assert bench.knn_stream_model(None, q) is None
