"""bench.py contract on the CPU: the reference arm (oracle port on the host cores) prints one JSON
line with the keys the driver reads.  The GPU arm is exercised on the B200 box by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--batch", "256"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "converged_nmpc_solves_per_sec"
    assert line["unit"] == "solves/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["config"]["batch_per_gpu"] == 256 and line["config"]["restoration"] is True
    assert line["cpu_baseline"]["port"] == "oracle/cxx"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


import pytest


def test_clock_summary_of_one_and_several_gpus():
    """`clocks`: rank 0's GPU in the contract's keys, every GPU of the job under `per_gpu`, throttle
    reasons collected over all of them."""
    sys.path.insert(0, ROOT)
    import bench
    row = lambda mhz, cap, w, t: [str(mhz), "1965", "Not Active", "Not Active", "Not Active", cap, str(w), str(t)]
    one = bench.summarize_clocks([[row(1965, "Not Active", 350.5, 41)] for _ in range(4)])
    assert one["sm_mhz"] == 1965.0 and one["sm_max_mhz"] == 1965.0 and one["reasons"] == [] and one["samples"] == 4
    assert one["power_w"] == 350.5 and "per_gpu" not in one
    two = bench.summarize_clocks([[row(1965, "Not Active", 300, 40), row(1800, "Active", 700, 70)] for _ in range(3)])
    assert two["sm_mhz"] == 1965.0 and two["reasons"] == ["sw_power_cap"]
    assert [g["sm_mhz"] for g in two["per_gpu"]] == [1965.0, 1800.0] and two["per_gpu"][1]["reasons"] == ["sw_power_cap"]
    none = bench.summarize_clocks([])
    assert none["sm_mhz"] is None and none["reasons"] == ["nvidia-smi unavailable"]
    odd = bench.summarize_clocks([[["[N/A]", "[N/A]", "Not Active", "Not Active", "Not Active", "Not Active", "[N/A]", "40"]]])
    assert odd["sm_mhz"] is None and odd["reasons"] == []


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["cfg2", "cfg4"])
def test_gpu_arm_json_line(config):
    """The GPU arm prints one JSON line with the keys of the bench contract (small batch, no CPU leg)."""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--config", config, "--steps", "2", "--warmup", "1",
           "--batch", "2048", "--inner", "4", "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert line["metric"] == "converged_nmpc_solves_per_sec" and line["unit"] == "solves/s" and line["value"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1 and line["dtype"] == "f64"
    assert line["config"]["name"] == config and line["config"]["restoration"] is True
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] >= 2 and "sm_mhz" in line["clocks"] and "roofline" in line and "ranks" in line
    if config == "cfg2":
        r = line["roofline"]
        assert r["bound"] == "fp64" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert line["no_restoration"]["value"] > 0 and line["single_solve_latency"]["p50_ms"] > 0
        assert line["closed_loop_tick_latency"]["p50_ms"] > 0
        # value / e2e go through the two-deep pipeline; the one-launch-at-a-time figure sits beside them
        assert line["serial"]["value"] > 0 and line["detail"]["pipeline_results_bit_identical_to_serial"] is True
        assert line["value"] > 0.9 * line["serial"]["value"]
