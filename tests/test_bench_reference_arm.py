"""The reference arm of ``bench.py`` (``--impl reference``) runs on host cores only, so it is checked here: one
JSON line with the contract's keys, the reference's own ``Stitcher`` class behind it where the build output
``oracle/_ref`` can be made or found, and a silent exit 0 for the ranks other than 0."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1_3x720p",
                        "--steps", "2", "--warmup", "1", "--ref-panos-per-step", "1"] + list(args),
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-2000:]
    return r.stdout.decode().strip().splitlines()


@pytest.mark.parametrize("port", [False, True])
def test_reference_arm_line(port):
    from oracle import build_ref
    lines = _run({"MCS_BENCH_REFERENCE_PORT": "1"} if port else None)
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "panoramas_per_sec" and d["unit"] == "panoramas/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "u8"
    assert d["config"]["workload"] == "cfg1_3x720p" and d["config"]["panoramas_per_step"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "panoramas/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    have_ref = build_ref.available() or build_ref.prebuilt() is not None
    assert cb["kind"] == ("reference" if have_ref and not port else "port")


def test_other_ranks_print_nothing():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, ["--gpus", "2"]) == []


def test_gpu_arm_refuses_to_run_without_a_device():
    """No CPU fallback: without a CUDA device the GPU arm ends with an error instead of a number."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode != 0 and r.stdout.decode().strip() == ""
    assert "no CUDA device" in r.stderr.decode(errors="replace")
