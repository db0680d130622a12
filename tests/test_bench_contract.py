"""The reference arm of bench.py runs on host cores only, so its JSON contract is checked here on CPU."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--scale", "0.002"],
                       capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reads_deduped_per_sec" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C5")
    # the sample size is part of the config: the two arms never claim the same config on different input sizes
    assert d["config"]["sample_reads"] == 400000 and d["pairs_per_s"] > 0 and d["dist_calls_per_s"] > 0 and d["scaling"] == "strong"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--scale", "0.002"],
                       capture_output=True, text=True, timeout=600, cwd=REPO, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_sample_of_a_single_bucket_workload_is_bounded():
    """C4 is ONE bucket: the faithful Naive scan is quadratic in it, so the CPU sample is capped at 80 k reads whatever
    --cpu-sample-reads says (2 M reads there would run for hours and stall `bench.py --workload C4`)."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sys.path.insert(0, os.path.join(REPO, "umi-collapse-rs_b200"))
    from umigpu import synth
    args = types.SimpleNamespace(scale=1.0, cpu_sample_reads=2_000_000)
    s4 = bench.cpu_sample_scale(args, synth.CONFIGS["C4"])
    assert int(synth.CONFIGS["C4"]["n_reads"] * s4) == 80_000
    s5 = bench.cpu_sample_scale(args, synth.CONFIGS["C5"])
    assert int(round(synth.CONFIGS["C5"]["n_reads"] * s5)) == 2_000_000
