"""bench.py's host-side pieces that need no GPU: the algorithmic bytes per lookup are SURVEY.md section 8(d)'s
figures, the measured-peak lookup, the nvidia-smi clock parser, the argument defaults of the contract."""
import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def test_bytes_per_lookup_are_the_survey_figures(bench):
    # 8 (int64 index) + 16 (one hash slot) + row_bytes + 4 d (fp32 output), P = 1
    want = {16: {32: 152, 16: 120, 8: 104, 4: 96}, 36: {32: 312, 16: 240, 8: 204, 4: 186}, 64: {32: 536, 16: 408, 8: 344, 4: 312}}
    for d, by_prec in want.items():
        for prec, b in by_prec.items():
            assert bench.bytes_per_lookup(d, prec) == b, (d, prec)


def test_measured_peak_comes_from_the_driver_file_or_the_stated_fallback(bench, tmp_path, monkeypatch):
    peak, src = bench.measured_peak_hbm()
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        assert peak == float(json.load(open(p))["hbm_gbs"]) and "MEASURED_PEAKS.json" in src
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    peak, src = bench.measured_peak_hbm()
    assert peak == 6650.0 and "fallback" in src


def test_clock_sampler_parses_nvidia_smi_lines(bench):
    s = bench.ClockSampler(0)
    assert s.stop()["reasons"] == ["unsampled"]                  # never started
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.lines = ["0, 1965, 1965, 412.50, Not Active, Not Active, Not Active, Not Active",
               "0, 1950, 1965, 640.10, Not Active, Not Active, Not Active, Active",
               "0, 1935, 1965, 700.00, Not Active, Not Active, Not Active, Active",
               "garbage line"]
    c = s.stop()
    assert c["sm_mhz"] == 1950.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"] and c["samples"] == 3


def test_contract_defaults(bench, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse_args()
    assert a.gpus == 1 and a.impl == "ours" and a.warmup >= 3 and a.steps > 0 and a.policy == "evlfu" and a.shape == "kaggle"
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "7", "--warmup", "3"])
    a = bench.parse_args()
    assert (a.impl, a.gpus, a.steps, a.warmup) == ("reference", 2, 7, 3)


def test_reference_arm_other_ranks_exit_quietly(bench, monkeypatch, capsys):
    """Under torchrun only rank 0 runs the reference arm; the other ranks print nothing and return 0."""
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "2"])
    assert bench.main_reference(bench.parse_args()) == 0
    assert capsys.readouterr().out == ""
