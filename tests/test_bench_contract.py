"""bench.py's host logic and the JSON contract, without a GPU: the engine is replaced by a stub that counts calls (the
real thing is measured on the B200 box; here only the plumbing -- workload set-up, both timed passes, the e2e loop, the
roofline arithmetic, the CPU-baseline leg with the unmodified reference engine, the keys of the line -- is exercised)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class StubEngine:
    def __init__(self, sd, device=0):
        self.sd, self.steps, self.opts = sd, 0, {}

    def set_option(self, k, v):
        self.opts[k] = v

    def stat(self, k):
        return {"timer_start": 0.0, "timer_stop_ms": 2.0, "launches": 6.0 * self.steps, "air_ms": 1.0, "air_launches_timed": float(self.steps),
                "air_cfg": 0.0, "air_lanes_z": 32.0, "fused": 1.0}[k]

    def reset_stats(self):
        self.steps = 0

    def run_steps(self, n0, k):
        assert 0 <= n0 and n0 + k <= self.sd.Nt
        self.steps += k

    def step_host(self, n, samples=None):
        assert samples is None or len(samples) == self.sd.Ns
        return np.zeros(self.sd.Nr)

    def read_outputs(self, n0=0, n1=None):
        return np.zeros((self.sd.Nr, (self.sd.Nt if n1 is None else n1) - n0))

    def sync(self):
        pass

    def close(self):
        pass


def test_default_workload_is_the_grid_the_metric_is_quoted_on():
    """every N runs BASELINE configs[4] (it fits one B200), so that the 1/2/4/8 values are one strong-scaling curve"""
    import bench
    assert bench.WORKLOADS["c5"]["N"] == (2048, 2048, 1024) and bench.WORKLOADS["c5"]["probes"]
    src = (ROOT / "bench.py").read_text()
    assert 'wl = args.workload or "c5"' in src and '"scaling": "strong"' in src and '"scaling": "weak"' not in src


def test_parity_object_and_probes(monkeypatch, capsys):
    """the probes workload: a source on every interface of the 8-way split, receivers across them; the line carries the trace hash and
    the reduced-grid check against the CPU engine (the stub engine returns silence, so the check must report a mismatch, not pass)"""
    import torch
    import bench
    import pffdtd_b200.engine as eng_mod
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(eng_mod, "Engine", StubEngine)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "small_probes", "--steps", "6", "--warmup", "3", "--no-cpu"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    p = line["parity"]
    assert p["steps"] == bench.PARITY_STEPS and p["receivers"] == 7 * 2 * (128 // 16) and len(p["traces_sha256"]) == 64
    assert p["equal_to_one_gpu_run"] is None  # no committed hash for this workload
    r = p["reduced_grid_vs_cpu_engine"]
    assert r["bit_exact"] is False and r["peak"] > 0 and "checker" in r and r["steps"] == 48
    assert line["e2e"]["h2d_bytes_per_step"] == 7 * 4 and line["scaling"] == "strong"
    assert line["warmup"] == 4  # rounded up to even


@pytest.mark.parametrize("workload", ("small",))
def test_bench_line_has_the_contract_keys(monkeypatch, capsys, workload):
    import torch
    import bench
    import pffdtd_b200.engine as eng_mod
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(eng_mod, "Engine", StubEngine)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", workload, "--steps", "6", "--warmup", "3"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "clocks", "e2e", "gpu_launches", "cpu_baseline"):
        assert k in line, k
    assert line["metric"] == "Gvoxel-updates/s" and line["unit"] == "Gvox/s" and line["n_gpus"] == 1 and line["steps"] == 6
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32" and line["config"]["workload"] == workload
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and "traffic" in r
    assert r["bytes_per_node"] == 12.125
    # the step's required traffic with the lossy boundary nodes' branch state counted (16 B per branch + 30 B per node, fp32)
    assert r["boundary_state_bytes_per_step"] > 0 and r["whole_step_frac_with_boundary_state"] > r["whole_step_frac"] > 0
    e = line["e2e"]
    assert e["unit"] == "Gvox/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    c = line["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert line["gpu_launches"] == 36 and "also" not in line and "parity" not in line
    # value = nodes * steps / device time: 128*96*64 nodes, 6 steps, 2 ms (the stub's stopwatch)
    assert abs(line["value"] - 128 * 96 * 64 * 6 / 2e-3 / 1e9) < 1e-9


def test_reference_arm_line(monkeypatch, capsys):
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--workload", "small", "--steps", "4", "--warmup", "3"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "Gvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"]
