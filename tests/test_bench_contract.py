"""bench.py's JSON contract, checked without a GPU.

bench.py talks to libplb through the C ABI only, so the host-side SIMT
emulation (tests/emu, test infrastructure) can stand in for the GPU at a tiny
scale: the numbers mean nothing, the KEYS and their types are what the driver
and the judge read.  `--impl reference` runs the CPU arm for real (tiny sample).
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))
import build_emu  # noqa: E402

BASE_KEYS = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int,
             "warmup": int, "ms_per_step": float, "higher_is_better": bool,
             "scaling": str, "dtype": str, "data": str, "config": dict, "e2e": dict,
             "gpu_launches": int, "cpu_baseline": dict}


def run_bench(*args, emulated=True):
    env = dict(os.environ)
    if emulated:
        env["PLB_LIB"] = build_emu.build()
    # bench.py itself only loads sm_100a builds; the emulated arm goes through
    # the harness wrapper, which flips the loader's test-only switch first
    head = ([os.path.join(HERE, "emu", "run_emulated.py")] if emulated else [])
    proc = subprocess.run([sys.executable] + head +
                          [os.path.join(REPO, "bench.py")] + list(args),
                          env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, proc.stdout            # exactly ONE JSON line
    return json.loads(lines[0])


def check_base(line):
    for key, typ in BASE_KEYS.items():
        assert key in line, key
        assert isinstance(line[key], typ), (key, line[key])
    assert "vs_baseline" in line and line["vs_baseline"] is None   # nothing published
    assert line["unit"] == "GLUPS" and line["dtype"] == "f64"
    assert line["higher_is_better"] is True and line["data"] == "synthetic"
    assert isinstance(line["config"]["workload"], str)
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in line["cpu_baseline"], key
    assert line["cpu_baseline"]["kind"] == "port"


def test_own_arm_line():
    line = run_bench("--scale", "0.02", "--steps", "11", "--warmup", "3",
                     "--cpu-sample", "128")
    check_base(line)
    assert line["n_gpus"] == 1 and line["steps"] == 11 and line["warmup"] == 3
    assert line["scaling"] == "weak"
    assert "MRT" in line["config"]["workload"] and "channel" in line["config"]["workload"]
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roof, key
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    # eleven steps of the MRT channel = two four-step passes + one three-step
    # pass, all counted
    reps = line["repeats"]
    assert reps >= 1 and len(line["repeat_ms"]) == reps
    assert roof["quads"] == 2 * reps and roof["triples"] == reps
    assert roof["pairs"] == 0 and roof["single_steps"] == 0
    assert roof["steps_per_launch"] == 4 and line["config"]["steps_per_pass"] == 4
    assert roof["frac"] <= roof["step_equivalent_frac"]
    assert "fused: block=" in roof["kernel_build"]
    assert line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] != line["value"]
    assert abs(line["e2e"]["global_mean_density"] - 1.0) < 1e-9    # mass is conserved
    assert "clocks" in line                                  # null without nvidia-smi
    # the benchmarked run checks itself against the oracle (device-resident
    # region and the end-to-end run)
    par = line["parity"]
    assert par["ok"] is True and par["max_rel_err"] <= 1e-12
    assert par["e2e_max_rel_err"] <= 1e-12
    assert par["steps_compared"] == 3 + 11 * reps + 1
    # configs[3] rides along, with its own parity
    assert "cavity" in line["extra"] and "cavity_16384_bgk" in line
    assert line["extra"]["cavity"]["parity"]["ok"] is True
    assert line["extra"]["cavity"]["scaling"] == "strong"


def test_two_steps_per_pass_are_accounted_for():
    env_key = "PLB_FUSE_DEPTH"
    old = os.environ.get(env_key)
    os.environ[env_key] = "2"
    try:
        line = run_bench("--scale", "0.02", "--steps", "9", "--warmup", "3",
                         "--no-extras", "--no-cpu-baseline")
    finally:
        if old is None:
            del os.environ[env_key]
        else:
            os.environ[env_key] = old
    roof = line["roofline"]
    reps = line["repeats"]
    assert roof["pairs"] == 4 * reps and roof["single_steps"] == reps and roof["triples"] == 0
    assert line["parity"]["ok"] is True
    assert roof["steps_per_launch"] == 2 and line["config"]["steps_per_pass"] == 2
    assert line["cpu_baseline"] is None


def test_reference_arm_line():
    line = run_bench("--impl", "reference", "--steps", "3", "--warmup", "3",
                     "--cpu-sample", "128", emulated=False)
    check_base(line)
    assert line["impl"] == "reference"
    assert line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "GLUPS",
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["value"] == line["value"]
