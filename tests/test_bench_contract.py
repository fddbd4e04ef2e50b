"""bench.py prints exactly one JSON line with the keys the driver's contract names."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline", "impl"}


def run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_contract():
    d = run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-sample", "384"])
    assert COMMON <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "sinkhorn_iters_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "128"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_contract_small():
    d = run(["--n", "6000", "--m", "5000", "--steps", "3", "--warmup", "3", "--cpu-sample", "384", "--no-aux"])
    assert COMMON <= set(d) and {"roofline", "clocks"} <= set(d)
    assert d["impl"] == "spadot_b200" and d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["config"]["finite"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1.5
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.2
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


def test_committed_bench_line_of_the_round_carries_the_contract():
    """profiles/r2_bench_tc_1Mx1M_final.json is what `bench.py --gpus 1 --steps 20 --warmup 5` printed on a B200 at the end of the
    round: the contract's keys, the roofline arithmetic (algorithmic exponentials / measured pass time) and the aux blocks."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_tc_1Mx1M_final.json")))
    assert COMMON <= set(d) and {"roofline", "clocks", "parity", "e2e_full_solve_s"} <= set(d)
    assert d["impl"] == "spadot_b200" and d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 5
    assert d["value"] == pytest.approx(1e3 / d["ms_per_step"], rel=1e-9)
    r = d["roofline"]
    assert r["bound"] == "sfu" and r["unit"] == "Tex2/s"
    assert r["achieved"] == pytest.approx(r["pairs_per_launch"] / (r["pass_ms_avg"] * 1e-3) / 1e12, rel=1e-6)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0.7 < r["frac"] < 0.85
    assert 2 * r["pass_ms_avg"] <= d["ms_per_step"] * 1.001                  # two passes per iteration fit inside a step
    assert d["e2e"]["value"] <= d["value"] and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["parity"]["lse_max_abs_err"] < 2e-5 and d["cpu_baseline"]["kind"] == "reference"
    assert "sw_thermal_slowdown" not in d["clocks"]["reasons"] and "hw_slowdown" not in d["clocks"]["reasons"]
    assert 40.0 < d["aux_syn_t_step"]["train_step_ms_median"] < 70.0 and d["aux_train_epoch"]["loss_trace_max_rel_diff"] < 1e-6
