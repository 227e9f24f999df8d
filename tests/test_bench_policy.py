"""CPU tests of the benchmark's headline policy (bench.choose_headline, bench.drift_stats): a mode that is not
bit-exact may only carry the headline when its MEASURED count of instances above the closed-loop bar is zero."""
import numpy as np
import pytest

import bench


def test_fast_mode_with_any_instance_above_the_bar_never_carries_the_headline():
    res = {"fast": (1.0, 9), "pipelined_exact": (1.7, 0), "onchip_exact": (2.2, 0)}
    assert bench.choose_headline(res, "pipelined_exact") == "pipelined_exact"


def test_fastest_green_mode_wins():
    res = {"fast": (1.0, 0), "pipelined_exact": (1.7, 0), "onchip_exact": (2.2, 0)}
    assert bench.choose_headline(res, "pipelined_exact") == "fast"
    res = {"pipelined_exact": (2.5, 0), "onchip_exact": (2.2, 0)}
    assert bench.choose_headline(res, "pipelined_exact") == "onchip_exact"


def test_without_an_anchor_only_bit_exact_modes_qualify():
    assert bench.choose_headline({"fast": (1.0, 0), "exact": (3.0, 0)}, None) == "exact"
    with pytest.raises(SystemExit):
        bench.choose_headline({"fast": (1.0, 0)}, None)


def test_drift_stats_counts_strictly_above_the_bar():
    d = np.array([0.0, 5e-7, 1e-6, 1.0000001e-6, 3e-6])
    st = bench.drift_stats(d)
    assert st["n_above_bar"] == 2 and st["instances"] == 5
    assert st["max_abs_dx"] == 3e-6
    x = np.zeros((4, 3))
    y = x.copy()
    y[2, 1] = 2e-6
    assert bench.parity_stats(y, x)["n_above_bar"] == 1


def test_candidate_tables_are_consistent():
    for model, cand in bench.CANDIDATES.items():
        assert model in bench.MODELS and model in bench.MODEL_ROW_DOUBLES
        assert any(m in bench.BIT_EXACT_MODES for m in cand), "every model needs a bit-exact anchor"
        assert all(m in bench.MODE_IDS and m in bench.KERNEL_OF_MODE for m in cand)
