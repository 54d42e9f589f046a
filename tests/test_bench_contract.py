"""bench.py's reference arm and JSON contract, on the CPU with a tiny workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--n-env", "4", "--n-dot", "4", "--res", "16", "--cpu-scans", "6"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ground_state_pixels_per_s" and d["unit"] == "pixels/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_flop_models_match_the_survey_figures():
    sys.path.insert(0, ROOT)
    import bench
    assert [bench.flops_reference_formulation(n) for n in (2, 4, 6, 8)] == [402.0, 1450.0, 6850.0, 39210.0]
    assert bench.flops_factored(8) < bench.flops_reference_formulation(8) / 30


def test_tunnel_cpu_baseline_block():
    """The CPU comparator of bench.py's tunnel-path block runs on its own (no GPU) and reports the contract's keys."""
    import bench
    from qdsim import synth
    dev = synth.sample_barrier_devices(2, 4, seed=3)
    mb = synth.tunnel_batch(dev)
    scans = synth.env_step_scans(mb, dev, res=8, seed=4)
    out = bench.tunnel_cpu_baseline(mb, scans, budget_s=0.2)
    assert out["kind"] == "port" and out["unit"] == "pixels/s" and out["value"] > 0 and out["cores"] >= 1
