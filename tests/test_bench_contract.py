"""The reference arm of bench.py (CPU only: the compiled reference on the host cores) prints the
contract's JSON line; the B200 arm refuses to run without a GPU (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(ref):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "1", "--cpu-procs", "2"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "classifiers/min" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_b200_arm_fails_loudly_without_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)


def _stub_detail(steps=20, warmup=5, lanes=24, world=1):
    """the B200 arm's full record as run_b200_arm assembles it, from stub statistics"""
    sys.path.insert(0, ROOT)
    import bench
    long = "x" * 900
    roof = {"bound": "popc", "kernel": "cell_gather_kernel", "achieved": 2251.123456789, "peak": 4498.5243, "unit": "Gpopc32/s",
            "frac": 0.50041234, "traffic": 263828480, "avg_launch_ms": 0.6597, "launches": 12961 * steps,
            "pair_evals_per_s": 1.5e12, "frac_reference_formulation": 1.33, "fp64_frac": 0.24, "note": long,
            "peak_source": long, "traffic_note": long, "timing": long,
            "in_bag_launches": {"frac": 0.6, "avg_launch_ms": 1.2, "note": long},
            "out_of_bag_launches": {"frac": 0.3, "avg_launch_ms": 0.2},
            "alone": {"frac": 0.47, "in_bag_frac": 0.6, "note": long},
            "screening": {"executed_fraction": 0.135, "effective_frac_reference_formulation": 3.1, "note": long},
            "em": {"frac": 0.4, "achieved": 1.0, "peak": 2.5, "unit": "Gadd/s", "sm_time_share": 0.28, "note": long},
            "sm_time": {"scoring_share": 0.5, "em_share": 0.3, "other_share": 0.2, "busy": 0.8},
            "frac_launch_events": 0.23, "frac_of_held_sm_time_in_bag": 0.66,
            "launch_events": {"frac": 0.23, "avg_launch_ms": 1.0, "note": long}}
    return {
        "metric": "classifiers/min trained (HLA-A 5k x 500 SNP)", "value": 861.123, "unit": "classifiers/min",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": 1672.5, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench.workload_config(world, lanes),
        "e2e": {"value": 870.0, "unit": "classifiers/min", "h2d_bytes_per_step": 244312994, "d2h_bytes_per_step": 164898565,
                "api": long},
        "e2e_legacy_hooks": {"value": 38.8, "api": long}, "gpu_launches": 107780 * steps,
        "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 78, "power_w_max": 667.15},
        "roofline": roof, "roofline_unscreened": {"frac": 0.65, "note": long},
        "cpu_baseline": {"value": 1.94, "unit": "classifiers/min", "cores": 16, "kind": "reference", "target": "64-bit, AVX512VPOPCNTDQ",
                         "sample": long, "calibration": long, "predict_value": 400.0, "predict_unit": "samples/s",
                         "parity_prefix_ok": True, "parity_prefix_snps": 13, "per_process": [{"a": 1}] * 16},
        "predict": {"value": 137000.0, "unit": "samples/s", "e2e": {"value": 90000.0}, "roofline": {"frac": 0.73},
                    "sharded_by_classifier_value": 1e5, "allreduce_ms": 3.0, "allreduce_bytes": 1316800000},
        "bed_decode": {"ms": 0.227, "note": long}, "large_list": {"frac": 0.58, "note": long},
        "train_detail": {"classifiers": {"count": (steps + warmup) * lanes}, "note": long},
        "device": {"name": "NVIDIA B200", "sm_count": 148, "clock_khz": 1965000},
    }


@pytest.mark.parametrize("steps,warmup,lanes,world", [(20, 5, 40, 1), (20, 5, 40, 8), (200, 50, 64, 8)])
def test_b200_line_is_bounded_and_complete(steps, warmup, lanes, world):
    """The driver parses ONE stdout line; round 1's grew with steps x lanes and was cut. The line is
    < 4 KB whatever the step count and carries the contract's keys as flat scalars."""
    sys.path.insert(0, ROOT)
    import bench
    line = bench.compact_line(_stub_detail(steps, warmup, lanes, world))
    text = json.dumps(line)
    assert len(text) < 4096, len(text)
    back = json.loads(text)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in back, key
    assert back["value"] > 0 and back["steps"] == steps and back["warmup"] == warmup
    assert back["config"]["lanes"] == lanes and "workload" in back["config"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "timing", "frac_launch_events", "alone_frac",
                "em_frac", "sm_time_busy", "global_operand_frac", "predict_frac"):
        assert key in back["roofline"], key
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in back["cpu_baseline"], key
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "predict_value"):
        assert key in back["e2e"], key
    # nothing nested survives below the contract's objects (the driver keeps flat members only)
    for obj in ("config", "e2e", "roofline", "cpu_baseline"):
        assert all(not isinstance(v, dict) for v in back[obj].values()), obj


def test_lanes_per_gpu_do_not_depend_on_the_number_of_gpus():
    sys.path.insert(0, ROOT)
    import bench
    assert len({bench.default_lanes(w) for w in (1, 2, 4, 8)}) == 1
