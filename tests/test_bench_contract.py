"""bench.py contract on the CPU: the reference arm prints exactly one JSON line with the required keys (it times the
oracle port; ~15 s here), and the work model / config table are self-consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "triples/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["steps"] == 1 and j["warmup"] == 0 and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "Freebase" in j["config"]["workload"]


def test_work_model_matches_survey_numerators():
    sys.path.insert(0, ROOT)
    import bench
    g, b_alg, f_alg, P = bench.work_model(dict(d=100, k=3, R=13, Ne=75043, Nw=67447), 20000, 200000, 105000, 105000)
    assert P == 7142578                                            # Freebase-shaped theta
    assert abs(f_alg - 4.08e9) < 1e7                               # F_alg = 6 d^2 k N_u + 8 d k N
    assert 1.7e8 < b_alg < 1.9e8                                   # B_alg ~ 178 MB
    assert set(g) == {"entity_build", "w_prep", "gemm_fwd", "score", "gemm_dw", "gemm_ge", "entity_pull", "word_pull", "finalize"}
    assert all(v["bytes"] > 0 for v in g.values())
    assert set(bench.CONFIGS) == {"c1", "c3", "c4", "c5"}
