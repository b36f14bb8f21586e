"""bench.py's output contract, checked where no GPU is needed: the reference arm (`--impl reference`) times the
reference's own DSK binary (oracle/_ref/dsk, else the oracle port) on the host cores and prints ONE JSON line with the
keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-bases", "120000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600,
                         env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    lines = [l for l in out.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Gbases/s" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["sample_bases"] == 120000 and d["config"]["workload"].startswith("C2:")


def test_other_ranks_of_the_reference_arm_do_nothing():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120,
                         env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.decode().strip() == ""


def test_clock_sampler_without_a_gpu_reports_no_samples():
    """The sampler opens NVML in its constructor (before the timed region); on a box without a GPU it must neither
    raise nor hang, and its summary keeps the keys of the bench line's `clocks` object."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.start()
    s.stop_flag = True
    s.join(timeout=10)
    assert not s.is_alive()
    c = s.summary()
    assert set(c) == {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
    assert c["samples"] == len(s.samples) and (c["samples"] > 0 or c["sm_mhz"] is None)
