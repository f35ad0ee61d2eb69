"""bench.py's driver contract, CPU side: the reference arm runs without a GPU and prints exactly one
JSON line with the keys the driver reads; the GPU arm's source carries every required key."""
import json
import os
import subprocess
import sys

import _libs


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OPENBLAS_NUM_THREADS="8")
    r = subprocess.run([sys.executable, os.path.join(_libs.ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "4096" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(_libs.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_emits_contract_keys():
    src = open(os.path.join(_libs.ROOT, "bench.py")).read()
    for key in ['"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"workload"', '"roofline"', '"bound"',
                '"achieved"', '"peak"', '"frac"', '"traffic"', '"cpu_baseline"', '"e2e"', '"h2d_bytes_per_step"',
                '"d2h_bytes_per_step"', '"gpu_launches"', '"clocks"', '"sm_mhz"', '"sm_max_mhz"', '"reasons"']:
        assert key in src, key
