"""CPU: the host-side pieces of bench.py that do not need a GPU -- the clock sampler degrades to "no sample" instead of
failing, the reference arm answers from a non-zero rank without work, the argument parser keeps the driver's contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_clock_sampler_without_a_gpu():
    import bench
    s = bench.ClockSampler(0)
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert out["sm_mhz"] is None or out["sm_mhz"] > 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_parser_contract():
    import bench
    argv = sys.argv
    try:
        sys.argv = ["bench.py", "--gpus", "1", "--steps", "7", "--warmup", "4"]
        a = bench.parse()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.steps == 7 and a.warmup == 4 and a.impl == "ours" and a.config == 4


def test_committed_bench_lines_carry_the_contract_keys():
    """profiles/r2_bench_ours.json is the line the last GPU visit printed: one JSON object with every key of the contract"""
    with open(os.path.join(ROOT, "profiles", "r2_bench_ours.json")) as fh:
        d = json.loads(fh.readline())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["peak"] > 0 and 0 < d["roofline"]["frac_exact"] < 1
    assert d["e2e"]["h2d_bytes_per_step"] == 640 * 480 * 4 and d["e2e"]["d2h_bytes_per_step"] == 640 * 480 * 5
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] > 0 and d["dtype"] == "f32" and d["vs_baseline"] is None
