"""bench.py's contract, as far as it can be checked without a GPU: the config table against SURVEY §8(d)'s per-unit
figures, the stale-counter refusal, and the reference arm's JSON line (config A: the reference's own CPU-runnable case)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config_table_matches_the_survey():
    import bench

    # decode bytes/frame = J*H*W*4 (+ J*12 + 16)
    assert bench.decode_bytes_per_frame(bench.CONFIGS["B"]) == 180_224 + 11 * 12 + 16
    assert bench.decode_bytes_per_frame(bench.CONFIGS["C"]) == 470_016 + 17 * 12 + 16
    assert bench.decode_bytes_per_frame(dict(bench.CONFIGS["C"], J=24)) == 663_552 + 24 * 12 + 16
    assert bench.decode_bytes_per_frame(bench.CONFIGS["D"]) == 720_896 + 11 * 12 + 16
    assert bench.canonical_hyp_flops(bench.CONFIGS["B"]) == 126_400 + 54 * 11
    assert bench.CONFIGS["B"]["frames"] == 4096 and bench.CONFIGS["B"]["H"] == 256 and bench.CONFIGS["B"]["scaling"] == "weak"
    assert bench.CONFIGS["D"]["frames"] == 65536 and bench.CONFIGS["D"]["H"] == 1024 and bench.CONFIGS["D"]["scaling"] == "strong"
    assert bench.CONFIGS["C"]["frames"] == 16384 and bench.ITERATIONS == 10000
    for world in (1, 2, 4, 8):
        assert (bench.CONFIGS["D"]["frames"] // world) % bench.CHUNK == 0  # strong scaling keeps whole chunks per rank


def test_stale_ncu_counters_are_refused(tmp_path, monkeypatch):
    import bench

    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "profiles")
    entry, why = bench.ncu_counters("B", "score")
    assert entry is None and "missing" in why
    hashes = {k: bench.kernel_source_hash(k) for k in bench.KERNEL_SOURCES}
    good = {"kernel_source_hashes": hashes, "commit": "abc1234", "configs": {"B": {"score_warp_instructions_per_launch": 1.0}}}
    (tmp_path / "profiles" / "ncu_counters.json").write_text(json.dumps(good))
    for kernel in ("score", "decode"):
        entry, why = bench.ncu_counters("B", kernel)
        assert why is None and entry["captured_at_commit"] == "abc1234"
    entry, why = bench.ncu_counters("D", "score")
    assert entry is None and "no entry" in why
    # each quoted kernel is stamped on its own: a change of the decode sources leaves the scoring counters valid
    (tmp_path / "profiles" / "ncu_counters.json").write_text(json.dumps(dict(good, kernel_source_hashes=dict(hashes, decode="0" * 16))))
    entry, why = bench.ncu_counters("B", "decode")
    assert entry is None and "stale" in why
    entry, why = bench.ncu_counters("B", "score")
    assert why is None


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "A", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["config"] == "A" and d["config"]["hypotheses"] == 256 and d["gpu_launches"] == 0


def test_static_bank_model_runs_on_the_shipped_library():
    """tools/sass_bank_model.py (profiles/hyp_bank_r2.md): the FP32 hypothesis kernel's issue-rate ceiling from its operand reads."""
    import shutil

    lib = os.path.join(ROOT, "spacecraft-pose-estimation_b200", "spe_b200", "libspe_b200.so")
    if shutil.which("cuobjdump") is None or not os.path.exists(lib):
        import pytest

        pytest.skip("needs cuobjdump and the built library")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_bank_model.py"), lib, "hypothesis_kernel_t1", "--weights=6,5,4,11"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    import re

    m = re.search(r"issue ceiling ([0-9.]+); at the measured ([0-9.]+) extra cycles -> ([0-9.]+)", res.stdout)
    assert m, res.stdout
    assert 0.6 < float(m.group(1)) < float(m.group(3)) < 0.9
