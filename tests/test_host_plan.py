"""Host side of the product without a GPU: the C-ABI library loads and exports what
include/srsce.h declares, and the planner's data-independent tables (raised-cosine taps, linear
interpolation rows, the CNN inpainting operator) agree with the oracle."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ce_oracle
from srsran_ce_pytorch_b200 import _lib, estimator, synth

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "srsce.h").read_text()
    declared = set(re.findall(r"\b(srsce_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.srsce_abi_version() == _lib.ABI_VERSION == 3


def _plan(pattern, n_prbs, n_layers, smoothing="filter", fill="linear", grid=52, syms=(2, 11), start=3, **kw):
    hop1 = synth.make_hop(grid, start, n_prbs, list(syms), pattern, (n_layers + 1) // 2)
    cfg = synth.make_config(30e3, smoothing, True, **kw)
    return estimator.make_plan(grid, 14, n_layers, hop1, synth.no_hop(), cfg, fill), hop1


@pytest.mark.parametrize("pattern", ["comb2", "comb4", "type2", "comb3", "comb6", "comb12", "dense"])
@pytest.mark.parametrize("n_prbs", [1, 2, 3, 7])
def test_filter_geometry_matches_oracle(pattern, n_prbs):
    plan, hop1 = _plan(pattern, n_prbs, 2)
    r = int(hop1.DMRSREmask[:, 0].sum())
    taps = ce_oracle.rc_taps(12 // r, min(3, n_prbs))
    got, _, _ = plan.tables()
    assert plan.info.n_taps == taps.size
    assert np.allclose(got, taps, rtol=0, atol=1e-15)
    assert plan.info.n_virtual == ce_oracle.n_virtual(taps.size, n_prbs, r)
    assert plan.info.n_re == n_prbs * r


@pytest.mark.parametrize("fill", ["linear", "cnn"])
@pytest.mark.parametrize("pattern", ["comb2", "comb4", "type2", "comb3", "comb6", "comb12", "dense"])
@pytest.mark.parametrize("n_prbs", [1, 2, 5, 24])
def test_fill_operator_matches_oracle(fill, pattern, n_prbs):
    """out = A . pilots with the planner's operator rows == oracle fill of the same pilots."""
    n_layers = 4
    plan, hop1 = _plan(pattern, n_prbs, n_layers, fill=fill)
    _, idx, coef = plan.tables()
    spec = ce_oracle.hop_from_any(hop1)
    rng = np.random.default_rng(7)
    n = plan.info.n_re
    for g in range(2):
        h = (rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))).astype(np.complex64)
        want = ce_oracle._fill_band(h, spec, g, "cnn" if fill == "cnn" else "tensorized")
        terms = coef.shape[2]
        got = np.zeros_like(want, dtype=np.complex128)
        for t in range(terms):
            got += coef[g, :, t, None].astype(np.float64) * h[np.minimum(idx[g] + t, n - 1)].astype(np.complex128)
        assert (idx[g] >= 0).all() and (idx[g] + terms <= max(n, terms)).all()
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < 2e-6, (fill, pattern, n_prbs, g, err)


def test_cnn_operator_full_band_is_narrow():
    plan, _ = _plan("comb4", 273, 4, fill="cnn", grid=273, syms=(13,), start=0)
    assert plan.info.fill_terms <= 4
    plan, _ = _plan("comb2", 273, 4, fill="cnn", grid=273, syms=(13,), start=0)
    assert plan.info.fill_terms <= 4


def test_plan_info_config2():
    plan, _ = _plan("comb4", 273, 4, grid=273, syms=(13,), start=0)
    i = plan.info
    assert (i.n_sc, i.n_re, i.n_cdm, i.n_taps, i.n_virtual, i.fft_size, i.cfo_defined) == (3276, 819, 2, 7, 3, 1024, 0)
    assert i.bytes_per_unit == 8 * (2 * 819 + 819 * 4 + 3276 * 14 * 4) + 48
    assert i.smem_bytes <= 227 * 1024


def test_error_behaviour_matches_reference():
    hop1 = synth.make_hop(52, 3, 4, [2, 11], "comb2", 1)
    cfg = synth.make_config()
    cfg.Smoothing = "bogus"
    with pytest.raises(ValueError, match="Unknown smoothing"):           # ce_rule_tensorized.py:668
        estimator.make_plan(52, 14, 1, hop1, synth.no_hop(), cfg)
    cfg = synth.make_config()
    cfg.CyclicPrefixDurations = cfg.CyclicPrefixDurations[:10]
    with pytest.raises(ValueError, match="length >= 14"):                # :815-816
        estimator.make_plan(52, 14, 1, hop1, synth.no_hop(), cfg)
    cfg = synth.make_config()
    hop2 = synth.make_hop(52, 20, 4, [11], "comb2", 1)
    with pytest.raises(AssertionError, match="overlap"):                 # :862
        estimator.make_plan(52, 14, 1, hop1, hop2, cfg)
    hop2 = synth.make_hop(52, 20, 4, [5], "comb4", 1)
    with pytest.raises(AssertionError, match="DM-RS mask"):              # :869
        estimator.make_plan(52, 14, 1, hop1, hop2, cfg)
    with pytest.raises(NotImplementedError):                             # n_sc > 4096: the TA IFFT would truncate
        estimator.make_plan(400, 14, 1, synth.make_hop(400, 0, 4, [2], "comb2", 1), synth.no_hop(), cfg)
    short = synth.make_hop(52, 3, 4, [0, 1], "comb2", 1, n_sym=2)
    with pytest.raises(IndexError):                                      # (n_sc, 2) grid + CFO on
        estimator.make_plan(52, 2, 1, short, synth.no_hop(), cfg)
    bad = synth.make_hop(52, 3, 4, [2, 11], "comb2", 1)
    bad.maskPRBs[30] = True
    with pytest.raises(ValueError, match="contiguous"):
        estimator.make_plan(52, 14, 1, bad, synth.no_hop(), cfg)


def test_no_cpu_fallback():
    """CPU tensors are refused; a plan made without a device refuses to compute."""
    hop1 = synth.make_hop(25, 4, 6, [2, 11], "comb2", 1)
    cfg = synth.make_config()
    data = synth.make_units(hop1, synth.no_hop(), cfg, 1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        estimator.srs_channel_estimator(data["rx"][0], data["pilots"][0], 1.0, hop1, synth.no_hop(), cfg)
    if not torch.cuda.is_available():
        from srsran_ce_pytorch_b200.compat import ce_rule_tensorized
        with pytest.raises(RuntimeError, match="no CPU fallback"):          # compat modules: CPU tensors refused by default
            ce_rule_tensorized.srs_channel_estimator(data["rx"][0], data["pilots"][0], 1.0, hop1, synth.no_hop(), cfg)
        estimator.accept_host_arrays(True)                                   # opt-in: host arrays are copied to a GPU --
        try:                                                                 # without one the call still raises
            with pytest.raises(RuntimeError, match="no CPU compute path"):
                ce_rule_tensorized.srs_channel_estimator(data["rx"][0], data["pilots"][0], 1.0, hop1, synth.no_hop(), cfg)
        finally:
            estimator.accept_host_arrays(False)
        plan = estimator.make_plan(25, 14, 1, hop1, synth.no_hop(), cfg)
        assert plan.info.has_device == 0
        import ctypes as C
        strides = (C.c_int64 * 3)(1, 1, 1)
        rc = _lib.lib.srsce_estimate(plan._h, 16, strides, 16, 1, 1.0, 16, 16, 1, None)
        assert rc == _lib.ERR_NO_DEVICE and "no CPU compute path" in _lib.last_error()


from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(bits0=st.integers(1, 4095), n_prbs=st.integers(1, 9), shift=st.integers(0, 11), fill=st.sampled_from(["linear", "cnn"]))
def test_fill_operator_property_random_masks(bits0, n_prbs, shift, fill):
    """Any 12-bit RE mask (second CDM group = a rotation of the first, so both select the same number of REs):
    the planner's operator rows reproduce the oracle's fill of random pilots."""
    col0 = "".join("1" if (bits0 >> k) & 1 else "0" for k in range(12))
    col1 = col0[shift:] + col0[:shift]
    hop1 = synth.make_hop(30, 2, n_prbs, [3], [col0, col1], 2)
    cfg = synth.make_config(30e3, "none", False)
    try:
        plan = estimator.make_plan(30, 14, 4, hop1, synth.no_hop(), cfg, fill)
    except NotImplementedError:
        return                                              # operator wider than the kernel's row budget: refused, not wrong
    _, idx, coef = plan.tables()
    spec = ce_oracle.hop_from_any(hop1)
    n = plan.info.n_re
    rng = np.random.default_rng(bits0)
    for g in range(2):
        h = (rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))).astype(np.complex64)
        want = ce_oracle._fill_band(h, spec, g, "cnn" if fill == "cnn" else "tensorized")
        got = np.zeros_like(want, dtype=np.complex128)
        for t in range(coef.shape[2]):
            got += coef[g, :, t, None].astype(np.float64) * h[np.minimum(idx[g] + t, n - 1)].astype(np.complex128)
        assert np.abs(got - want).max() <= 3e-6 * max(np.abs(want).max(), 1e-30)


@pytest.mark.parametrize("fill", ["linear", "cnn"])
@pytest.mark.parametrize("pattern,n_prbs,grid", [("comb4", 273, 273), ("comb2", 273, 273), ("comb4", 24, 52), ("comb2", 9, 52),
                                                   ("type2", 17, 52), ("comb3", 8, 52), ("comb6", 30, 52), ("comb12", 16, 52),
                                                   ("dense", 11, 52), ("comb4", 1, 52), ("comb2", 3, 52)])
def test_tensor_fill_tiles_reproduce_the_operator(fill, pattern, n_prbs, grid):
    """The bf16 operator tiles of the tensor-core fill, summed over their terms and placed at each tile's pilot window,
    are exactly the float32 operator rows of the direct path (which test_fill_operator_matches_oracle pins to the oracle)."""
    plan, hop1 = _plan(pattern, n_prbs, 4, fill=fill, grid=grid, start=grid - n_prbs)
    i = plan.info
    if not i.fill_tensor:
        pytest.skip("no tensor-core fill for this geometry")
    _, idx, coef = plan.tables()
    a, kind, k0 = plan.fill_tensor_tables()
    rows, terms = 12 * n_prbs, coef.shape[2]
    assert i.fill_tensor_tiles == -(-rows // 96) and i.fill_tensor_k % 16 == 0
    assert i.fill_tensor_terms == {1: 3, 2: 5, 3: 6}[i.fill_tensor_a_terms]
    for g in range(i.n_cdm):
        dense = np.zeros((rows, i.n_re + i.fill_tensor_k), dtype=np.float64)
        for r in range(rows):
            dense[r, idx[g, r]:idx[g, r] + terms] += coef[g, r].astype(np.float64)
        for m in range(i.fill_tensor_tiles):
            r0, r1 = 96 * m, min(rows, 96 * m + 96)
            got = np.zeros_like(dense[r0:r1])
            got[:, k0[m]:k0[m] + i.fill_tensor_k] = a[kind[m], g, :r1 - r0]
            assert np.array_equal(got, dense[r0:r1]), (g, m)
            assert not a[kind[m], g, r1 - r0:].any()
            assert not dense[r0:r1, i.n_re:].any()


def test_tensor_fill_is_refused_where_it_does_not_apply():
    hop1 = synth.make_hop(52, 2, 8, [1], "comb4", 2)
    hop2 = synth.make_hop(52, 30, 8, [8], "comb4", 2)
    plan = estimator.make_plan(52, 14, 4, hop1, hop2, synth.make_config(30e3, "filter", True))
    assert plan.info.fill_tensor == 0
    with pytest.raises(NotImplementedError, match="two hops"):
        plan.fill_tensor_tables()


def test_every_geometry_inside_the_limits_plans_and_the_global_scratch_envelope_is_the_documented_one():
    """inside n_sc <= 4096, <= 4 layers, <= 14 symbols every geometry plans.  The only plans whose
    smoothed pilots do not fit in shared memory (and live in a per-CTA global scratch: ``info.h_in_global``) are two-hop
    plans with the dense mask at 273 PRB with >= 3 layers and at 341 PRB with >= 2 layers."""
    in_global = set()
    for grid in (106, 273, 341):
        for pat in ("comb2", "comb4", "type2", "comb6", "dense"):
            for n_layers in (1, 2, 3, 4):
                for syms, syms2 in (([13], None), ([2, 5, 8, 11], None), ([3], [10]), ([2, 4], [9, 11])):
                    hop1 = synth.make_hop(grid, 0, grid, syms, pat, (n_layers + 1) // 2)
                    hop2 = synth.make_hop(grid, 0, grid, syms2, pat, (n_layers + 1) // 2) if syms2 else synth.no_hop()
                    plan = estimator.make_plan(grid, 14, n_layers, hop1, hop2, synth.make_config(30e3, "filter", True))
                    assert plan.info.smem_bytes <= 227 * 1024
                    if plan.info.h_in_global:
                        in_global.add((grid, pat, n_layers, syms2 is not None))
    want = {(273, "dense", 3, True), (273, "dense", 4, True), (341, "dense", 2, True), (341, "dense", 3, True), (341, "dense", 4, True)}
    assert in_global == want, in_global ^ want
