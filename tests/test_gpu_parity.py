"""Parity of the CUDA path (through the drop-in entry points, i.e. through the C ABI) with
  * outputs of the unmodified reference stored in tests/golden (all three variants), and
  * the numpy oracle on fresh seeded inputs, batched,
plus size-independent properties at BASELINE.json's full config-2 size.
Floating point: tolerances are in tests/tolerances.py (grid max-rel 1e-5 and NMSE 1e-12 gates)."""
import numpy as np
import pytest
import torch

from oracle import ce_oracle
from tests import tolerances as tol
from pathlib import Path

from tests.golden_io import case_names, load_case

ROOT = Path(__file__).resolve().parents[1]

pytestmark = pytest.mark.gpu
NAMES = case_names()


def _dev(x):
    return torch.as_tensor(x).cuda()


def _scalars(out):
    cfo = out[5]
    return [float(out[1]), float(out[2]), float(out[3]), float(out[4]), float(cfo) if cfo.numel() else None]


def _ta_margin(c, variant):
    diag = {}
    ce_oracle.estimate(c["rx"], c["pilots"], c["beta"], c["hop1"], c["hop2"], c["cfg"],
                       fill="tensorized" if variant != "cnn" else "tensorized", diagnostics=diag)
    return diag["ta_margin"]


@pytest.mark.parametrize("name", NAMES)
def test_rule_estimator_matches_reference(name):
    """ce_rule_tensorized / ce_rule_baseline drop-ins vs the stored reference outputs."""
    from srsran_ce_pytorch_b200.compat import ce_rule_baseline, ce_rule_tensorized
    c = load_case(name)
    margin = _ta_margin(c, "tensorized")
    for variant, mod in (("tensorized", ce_rule_tensorized), ("baseline", ce_rule_baseline)):
        if f"{variant}_grid" not in c["ref"]:
            continue
        out = mod.srs_channel_estimator(_dev(c["rx"]), _dev(c["pilots"]), c["beta"], c["hop1"], c["hop2"], c["cfg"])
        assert out[0].dtype == torch.complex64 and out[1].dtype == torch.float64 and out[0].is_cuda
        mx, nmse = tol.grid_errors(out[0].cpu().numpy(), c["ref"][f"{variant}_grid"])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (name, variant, mx, nmse)
        tol.check_scalars(_scalars(out), c["ref"][f"{variant}_scalars"], float(c["cfg"].scs), margin,
                          tag=f"{name}/{variant}")


def test_zero_code_switch_with_host_arrays():
    """with compat/ ahead of the reference's src/ on sys.path, ``from ce_rule_tensorized import
    ...`` works; after the opt-in ``accept_host_arrays()`` it serves the reference's own call shape -- HOST tensors (or
    numpy arrays) in, CPU tensors out, as scripts/validation/validate_all.py:367-382 / :543 / :549-561 use it -- with the
    arithmetic on the GPU.  Without the opt-in CPU tensors raise."""
    import importlib
    import sys
    compat_dir = str(ROOT / "srsran_ce_pytorch_b200" / "compat")
    sys.path.insert(0, compat_dir)
    try:
        mods = [importlib.import_module(n) for n in ("ce_rule_tensorized", "ce_rule_baseline", "ce_dl_cnn")]
    finally:
        sys.path.remove(compat_dir)
        for n in ("ce_rule_tensorized", "ce_rule_baseline", "ce_dl_cnn"):
            sys.modules.pop(n, None)
    from srsran_ce_pytorch_b200 import estimator
    c = load_case(NAMES[0])
    rx, pil = torch.as_tensor(c["rx"]).cpu(), torch.as_tensor(c["pilots"]).cpu()
    with pytest.raises(RuntimeError, match="no CPU fallback"):               # the default: CPU tensors are refused
        mods[0].srs_channel_estimator(rx, pil, c["beta"], c["hop1"], c["hop2"], c["cfg"])
    estimator.accept_host_arrays(True)
    try:
        _host_array_calls(mods, c, rx, pil)
    finally:
        estimator.accept_host_arrays(False)


def _host_array_calls(mods, c, rx, pil):
    for mod, variant in zip(mods, ("tensorized", "baseline", "cnn")):
        if f"{variant}_grid" not in c["ref"]:
            continue
        assert {"HopConfig", "EstimatorConfig", "srs_channel_estimator"} <= set(dir(mod))
        out = mod.srs_channel_estimator(rx, pil, c["beta"], c["hop1"], c["hop2"], c["cfg"])
        assert all(not t.is_cuda for t in out) and out[0].dtype == torch.complex64 and out[1].dtype == torch.float64
        mx, nmse = tol.grid_errors(out[0].numpy(), c["ref"][f"{variant}_grid"])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (variant, mx, nmse)
        tol.check_scalars(_scalars(out), c["ref"][f"{variant}_scalars"], float(c["cfg"].scs), _ta_margin(c, variant if variant == "cnn" else "tensorized"),
                          tag=f"host/{variant}")
        on_dev = mod.srs_channel_estimator(rx.cuda(), pil.cuda(), c["beta"], c["hop1"], c["hop2"], c["cfg"])
        assert on_dev[0].is_cuda and all(torch.equal(a, b.cpu()) for a, b in zip(out, on_dev))
        as_numpy = mod.srs_channel_estimator(rx.numpy(), pil.numpy(), c["beta"], c["hop1"], c["hop2"], c["cfg"])
        assert all(torch.equal(a, b) for a, b in zip(out, as_numpy))


@pytest.mark.parametrize("name", [n for n in NAMES])
def test_cnn_estimator_matches_reference(name):
    from srsran_ce_pytorch_b200.compat import ce_dl_cnn
    c = load_case(name)
    if "cnn_grid" not in c["ref"]:
        pytest.skip("no ce_dl_cnn output stored for this case")
    out = ce_dl_cnn.srs_channel_estimator(_dev(c["rx"]), _dev(c["pilots"]), c["beta"], c["hop1"], c["hop2"], c["cfg"])
    mx, nmse = tol.grid_errors(out[0].cpu().numpy(), c["ref"]["cnn_grid"])
    assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (name, mx, nmse)
    tol.check_scalars(_scalars(out), c["ref"]["cnn_scalars"], float(c["cfg"].scs), _ta_margin(c, "cnn"),
                      tag=f"{name}/cnn")


BATCH_CASES = [
    # grid, start, n_prbs, symbols, pattern, layers, smoothing, cfo, hop2 (start, symbols), pilot_share, symbol_major
    (52, 5, 12, [2, 11], "comb2", 2, "filter", True, None, 2, False),
    (52, 0, 52, [3], "comb4", 4, "filter", True, None, 4, True),
    (106, 10, 24, [2, 7, 11], "type2", 3, "filter", True, None, 1, False),
    (52, 2, 8, [1, 3], "comb4", 4, "mean", True, (30, [8, 10]), 2, True),
    (52, 2, 8, [2], "comb6", 2, "none", False, (40, [9]), 1, False),
    (25, 0, 25, [0, 5, 10], "dense", 1, "filter", False, None, 1, False),
    (52, 7, 3, [4, 9], "comb3", 2, "filter", True, None, 3, False),
    (52, 7, 2, [4], "comb12", 1, "filter", True, None, 1, False),
    (52, 7, 1, [4, 6], "comb6", 1, "filter", True, None, 1, False),
]


@pytest.mark.parametrize("case", BATCH_CASES, ids=[f"b{i}" for i in range(len(BATCH_CASES))])
@pytest.mark.parametrize("fill", ["linear", "cnn"])
def test_batched_units_match_oracle(case, fill):
    """batched[b] == oracle(rx[b], pilots[b // share]) for every unit of a seeded batch."""
    from srsran_ce_pytorch_b200 import estimator, synth
    grid, start, n, syms, pat, L, sm, cfo, h2, share, sym_major = case
    hop1 = synth.make_hop(grid, start, n, syms, pat, (L + 1) // 2)
    hop2 = synth.no_hop() if h2 is None else synth.make_hop(grid, h2[0], n, h2[1], pat, (L + 1) // 2)
    cfg = synth.make_config(30e3, sm, cfo)
    n_units = 6 * share
    data = synth.make_units(hop1, hop2, cfg, L, n_units, pilot_share=share, seed=99, device="cuda",
                            symbol_major=sym_major, beta=1.25)
    out, sc = estimator.srs_channel_estimator_batched(data["rx"], data["pilots"], data["beta"], hop1, hop2, cfg,
                                                      fill=fill, pilot_share=share)
    out, sc = out.cpu().numpy(), sc.cpu().numpy()
    rx, pil = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
    for b in range(n_units):
        diag = {}
        want = ce_oracle.estimate(rx[b], pil[b // share], data["beta"], hop1, hop2, cfg,
                                  fill="cnn" if fill == "cnn" else "tensorized", diagnostics=diag)
        mx, nmse = tol.grid_errors(out[b], want[0])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (case, fill, b, mx, nmse)
        got = [sc[b, 0], sc[b, 1], sc[b, 2], sc[b, 3], sc[b, 4] if sc[b, 5] else None]
        tol.check_scalars(got, want[1:], float(cfg.scs), diag["ta_margin"], tag=f"{case}/{fill}/{b}")


GLOBAL_H_CASES = [
    # grid PRBs, layers, hop-1 pilot symbols, hop-2 pilot symbols, smoothing, cfo compensation, fill
    (273, 4, [3], [10], "filter", True, "linear"),
    (273, 3, [2, 4], [9, 11], "filter", True, "linear"),       # CFO estimated and compensated: final rotation
    (341, 2, [3], [10], "mean", True, "linear"),
    (341, 4, [2, 4], [9, 11], "filter", False, "linear"),
    (341, 3, [3], [10], "none", True, "cnn"),
]


@pytest.mark.parametrize("case", GLOBAL_H_CASES, ids=[f"g{i}" for i in range(len(GLOBAL_H_CASES))])
def test_largest_two_hop_geometries_run_with_the_smoothed_pilots_in_global_memory(case):
    """The dense mask on two full-band hops: the smoothed pilots of both hops do not fit in shared memory, the planner
    keeps them in a per-CTA global scratch (``info.h_in_global``).  Same parity bar as every other geometry, for more
    units than there are resident CTAs (the scratch is reused across the units of a CTA), plus the fill seam."""
    from srsran_ce_pytorch_b200 import estimator, synth
    grid, L, syms1, syms2, sm, cfo, fill = case
    hop1 = synth.make_hop(grid, 0, grid, syms1, "dense", (L + 1) // 2, start_symbol=0, n_alloc=7)
    hop2 = synth.make_hop(grid, 0, grid, syms2, "dense", (L + 1) // 2, start_symbol=7, n_alloc=7)
    cfg = synth.make_config(30e3, sm, cfo)
    plan = estimator.make_plan(grid, 14, L, hop1, hop2, cfg, fill)
    assert plan.info.h_in_global == 1
    n_units = 2 * torch.cuda.get_device_properties(0).multi_processor_count + 3
    data = synth.make_units(hop1, hop2, cfg, L, n_units, seed=5, device="cuda")
    out, sc = plan.estimate(data["rx"], data["pilots"], 1.0, 1)
    again, sc_again = plan.estimate(data["rx"][-4:], data["pilots"][-4:], 1.0, 1)
    assert torch.equal(again, out[-4:]) and torch.equal(sc_again, sc[-4:])
    rx, pil = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
    o, s = out.cpu().numpy(), sc.cpu().numpy()
    for b in (0, n_units // 2, n_units - 1):
        diag = {}
        want = ce_oracle.estimate(rx[b], pil[b], 1.0, hop1, hop2, cfg, fill="cnn" if fill == "cnn" else "tensorized",
                                  diagnostics=diag)
        mx, nmse = tol.grid_errors(o[b], want[0])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.grid_nmse_gate([syms1, syms2], cfo, sm), (case, b, mx, nmse)
        got = [s[b, 0], s[b, 1], s[b, 2], s[b, 3], s[b, 4] if s[b, 5] else None]
        tol.check_scalars(got, want[1:], float(cfg.scs), diag["ta_margin"], tag=f"{case}/{b}")
    # compact output of the same plan rebuilds the dense grid
    ch, rot, sc2 = plan.estimate_compact(data["rx"][:8], data["pilots"][:8], 1.0, 1)
    mx, nmse = tol.grid_errors(plan.expand_compact(ch, rot).cpu().numpy(), o[:8])
    assert mx <= 1e-6 and nmse <= 1e-13 and torch.equal(sc2, sc[:8]), (mx, nmse)
    # the fill seam of the same plan (direct implementation; smoothed pilots in, no final rotation)
    i = plan.info
    g = torch.Generator(device="cuda").manual_seed(1)
    h = torch.view_as_complex(torch.randn((5, 2, i.n_re, L, 2), device="cuda", generator=g))
    filled = plan.fill(h, impl="direct")
    from tests.test_gpu_fill import _expected
    want = _expected(h[:2].cpu().numpy(), hop1, hop2, i.n_sc, 14, fill)
    mx, nmse = tol.grid_errors(filled[:2].cpu().numpy(), want)
    assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (mx, nmse)
    assert torch.isfinite(torch.view_as_real(filled)).all()


def _config2(n_units, syms=(13,), seed=1234, share=4):
    from srsran_ce_pytorch_b200 import synth
    hop1 = synth.make_hop(273, 0, 273, list(syms), "comb4", 2)
    cfg = synth.make_config(30e3, "filter", True)
    data = synth.make_units(hop1, synth.no_hop(), cfg, 4, n_units, pilot_share=share, seed=seed, device="cuda")
    return hop1, synth.no_hop(), cfg, data


def test_config2_full_size_properties():
    """BASELINE config 2 (273 PRB, 4 ports, 4 rx, comb-4, 64 UEs = 256 units): properties that need no
    oracle run -- batch independence (bit exact), homogeneity under a complex gain, and a spot check of
    a few units against the oracle."""
    from srsran_ce_pytorch_b200 import estimator
    hop1, hop2, cfg, data = _config2(256)
    out, sc = estimator.srs_channel_estimator_batched(data["rx"], data["pilots"], 1.0, hop1, hop2, cfg, pilot_share=4)
    assert out.shape == (256, 3276, 14, 4) and torch.isfinite(torch.view_as_real(out)).all()
    # batch independence: a sub-batch gives bit-identical results
    sub, sc_sub = estimator.srs_channel_estimator_batched(data["rx"][64:72], data["pilots"][16:18], 1.0, hop1, hop2,
                                                          cfg, pilot_share=4)
    assert torch.equal(sub, out[64:72]) and torch.equal(sc_sub, sc[64:72])
    # homogeneity: rx * g  ->  grid * g, rsrp/epre/noise * |g|^2, same time alignment
    g = complex(0.6, -1.7)
    out_g, sc_g = estimator.srs_channel_estimator_batched(data["rx"][:16] * g, data["pilots"][:4], 1.0, hop1, hop2,
                                                          cfg, pilot_share=4)
    mx, nmse = tol.grid_errors(out_g.cpu().numpy(), (out[:16] * g).cpu().numpy())
    assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (mx, nmse)
    assert torch.allclose(sc_g[:, :3], sc[:16, :3] * abs(g) ** 2, rtol=1e-5, atol=0)
    assert torch.equal(sc_g[:, 3], sc[:16, 3])
    # spot check against the oracle
    rx, pil = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
    for b in (0, 101, 255):
        diag = {}
        want = ce_oracle.estimate(rx[b], pil[b // 4], 1.0, hop1, hop2, cfg, diagnostics=diag)
        mx, nmse = tol.grid_errors(out[b].cpu().numpy(), want[0])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (b, mx, nmse)
        s = sc[b].cpu().numpy()
        tol.check_scalars([s[0], s[1], s[2], s[3], s[4] if s[5] else None], want[1:], 30e3, diag["ta_margin"], tag=f"cfg2/{b}")


def test_symbol_major_input_is_equivalent():
    """The batched entry point takes any strides: a [B, n_sym, n_sc] grid viewed as [B, n_sc, n_sym]
    gives bit-identical results to the reference layout."""
    from srsran_ce_pytorch_b200 import estimator
    hop1, hop2, cfg, data = _config2(8, syms=(3, 10))
    a, sa = estimator.srs_channel_estimator_batched(data["rx"], data["pilots"], 1.0, hop1, hop2, cfg, pilot_share=4)
    rx_t = data["rx"].transpose(1, 2).contiguous().transpose(1, 2)
    assert rx_t.stride() != data["rx"].stride()
    b, sb = estimator.srs_channel_estimator_batched(rx_t, data["pilots"], 1.0, hop1, hop2, cfg, pilot_share=4)
    assert torch.equal(a, b) and torch.equal(sa, sb)


def test_empty_batch_and_return_types():
    from srsran_ce_pytorch_b200 import estimator
    hop1, hop2, cfg, data = _config2(4)
    out, sc = estimator.srs_channel_estimator_batched(data["rx"][:0], data["pilots"][:0], 1.0, hop1, hop2, cfg,
                                                      pilot_share=4)
    assert out.shape == (0, 3276, 14, 4) and sc.shape == (0, 6)
    one = estimator.srs_channel_estimator(data["rx"][0], data["pilots"][0], 1.0, hop1, hop2, cfg)
    assert one[0].shape == (3276, 14, 4) and one[5].shape == (0,) and one[5].dtype == torch.float64
    assert all(t.dim() == 0 and t.dtype == torch.float64 for t in one[1:5])


def test_persistent_loop_many_units_per_cta():
    """More units than resident CTAs: every CTA walks several units through its double-buffered
    staging / estimator / writer pipeline.  Batch independence must still hold bit for bit."""
    from srsran_ce_pytorch_b200 import estimator, synth
    for two_hops in (False, True):
        hop1 = synth.make_hop(24, 2, 8, [1, 3], "comb4", 2)
        hop2 = synth.make_hop(24, 12, 8, [8, 10], "comb4", 2) if two_hops else synth.no_hop()
        cfg = synth.make_config(30e3, "filter", True)
        n = 3000
        data = synth.make_units(hop1, hop2, cfg, 4, n, pilot_share=4, seed=77, device="cuda")
        out, sc = estimator.srs_channel_estimator_batched(data["rx"], data["pilots"], 1.0, hop1, hop2, cfg, pilot_share=4)
        for b0 in (0, 1488, 2992):
            sub, sc_sub = estimator.srs_channel_estimator_batched(data["rx"][b0:b0 + 8], data["pilots"][b0 // 4:b0 // 4 + 2],
                                                                  1.0, hop1, hop2, cfg, pilot_share=4)
            assert torch.equal(sub, out[b0:b0 + 8]) and torch.equal(sc_sub, sc[b0:b0 + 8]), (two_hops, b0)
        rx, pil = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
        for b in (5, 1777, 2999):
            diag = {}
            want = ce_oracle.estimate(rx[b], pil[b // 4], 1.0, hop1, hop2, cfg, diagnostics=diag)
            mx, nmse = tol.grid_errors(out[b].cpu().numpy(), want[0])
            assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (two_hops, b, mx, nmse)
            s = sc[b].cpu().numpy()
            tol.check_scalars([s[0], s[1], s[2], s[3], s[4] if s[5] else None], want[1:], 30e3, diag["ta_margin"],
                              tag=f"persistent/{two_hops}/{b}")


@pytest.mark.parametrize("case", [BATCH_CASES[0], BATCH_CASES[3], BATCH_CASES[4], BATCH_CASES[6]], ids=["c0", "c3", "c4", "c6"])
@pytest.mark.parametrize("fill", ["linear", "cnn"])
def test_compact_output_expands_to_dense(case, fill):
    """srsce_estimate_compact: factors [B, hops, n_sc, L] x [B, n_sym] rebuild the dense grid
    of the reference contract; scalars are identical to the dense call."""
    from srsran_ce_pytorch_b200 import estimator, synth
    grid, start, n, syms, pat, L, sm, cfo, h2, share, sym_major = case
    hop1 = synth.make_hop(grid, start, n, syms, pat, (L + 1) // 2, start_symbol=1, n_alloc=12)
    hop2 = synth.no_hop() if h2 is None else synth.make_hop(grid, h2[0], n, h2[1], pat, (L + 1) // 2, start_symbol=5, n_alloc=9)
    cfg = synth.make_config(30e3, sm, cfo)
    data = synth.make_units(hop1, hop2, cfg, L, 4 * share, pilot_share=share, seed=11, device="cuda", symbol_major=sym_major)
    plan = estimator.make_plan(grid, 14, L, hop1, hop2, cfg, fill)
    dense, sc = plan.estimate(data["rx"], data["pilots"], 1.0, share)
    ch, rot, sc2 = plan.estimate_compact(data["rx"], data["pilots"], 1.0, share)
    assert ch.shape == (4 * share, 2 if h2 else 1, 12 * grid, L) and rot.shape == (4 * share, 14)
    assert torch.equal(sc, sc2)
    mx, nmse = tol.grid_errors(plan.expand_compact(ch, rot).cpu().numpy(), dense.cpu().numpy())
    assert mx <= 1e-6 and nmse <= 1e-13, (mx, nmse)


ODD_ROW_CASES = [
    # n_sym, layers, pilot symbols, start_symbol, n_alloc  -> row_elems = n_sym * layers
    (1, 1, [0], 0, 1),        # 1 element per row: 8-byte store path
    (1, 3, [0], 0, 1),        # 3 elements
    (7, 1, [3], 1, 5),        # 7 elements, partial allocation
    (13, 3, [6], 0, 13),      # 39 store slots per row: more than one warp pass per row
    (5, 3, [1, 3], 0, 5),     # two pilot symbols, CFO estimated but not compensated (grid is not 14 symbols)
    (14, 3, [2, 11], 0, 14),  # even row (42 elements) but odd layer count: unpaired slots, with the final CFO rotation
    (9, 2, [4], 2, 6),        # even row, 9 slots
]


@pytest.mark.parametrize("case", ODD_ROW_CASES, ids=[f"nsym{c[0]}-L{c[1]}" for c in ODD_ROW_CASES])
@pytest.mark.parametrize("two_hops", [False, True], ids=["1hop", "2hop"])
def test_short_and_odd_grids_match_oracle(case, two_hops):
    """Grids that are not 14 symbols / rows with an odd number of elements exercise the scalar (8-byte) store
    path and the multi-pass row loop of the writer warps (the reference allows them when CFO compensation does
    not apply)."""
    from srsran_ce_pytorch_b200 import estimator, synth
    n_sym, L, syms, start, n_alloc = case
    if two_hops and len(syms) < 2:
        if n_sym < 3:
            pytest.skip("a second hop needs a second pilot symbol")
        syms2 = [s for s in range(n_sym) if s not in syms][-1:]
    elif two_hops:
        syms, syms2 = syms[:1], syms[1:]
    cfo_comp = n_sym == 14
    hop1 = synth.make_hop(40, 3, 6, syms, "comb2", (L + 1) // 2, n_sym=n_sym, start_symbol=start, n_alloc=n_alloc)
    hop2 = synth.make_hop(40, 20, 6, syms2, "comb2", (L + 1) // 2, n_sym=n_sym, start_symbol=start, n_alloc=n_alloc) \
        if two_hops else synth.no_hop()
    cfg = synth.make_config(30e3, "filter", cfo_comp)
    data = synth.make_units(hop1, hop2, cfg, L, 5, pilot_share=1, seed=3, device="cuda")
    out, sc = estimator.srs_channel_estimator_batched(data["rx"], data["pilots"], 1.0, hop1, hop2, cfg, pilot_share=1)
    out, sc = out.cpu().numpy(), sc.cpu().numpy()
    rx, pil = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
    for b in range(5):
        diag = {}
        want = ce_oracle.estimate(rx[b], pil[b], 1.0, hop1, hop2, cfg, diagnostics=diag)
        mx, nmse = tol.grid_errors(out[b], want[0])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (case, two_hops, b, mx, nmse)
        got = [sc[b, 0], sc[b, 1], sc[b, 2], sc[b, 3], sc[b, 4] if sc[b, 5] else None]
        tol.check_scalars(got, want[1:], 30e3, diag["ta_margin"], tag=f"{case}/{two_hops}/{b}")


@pytest.mark.parametrize("pattern,syms,n_units,grid", [("comb4", [3, 10], 720, 273), ("comb2", [13], 400, 273),
                                                        ("comb2", [3, 10], 330, 273), ("type2", [6], 640, 273),
                                                        ("comb4", [2, 5, 8, 11], 340, 273), ("dense", [3, 10], 320, 273),
                                                        ("comb2", [1, 4, 8, 12], 320, 341)],
                         ids=["slim-2cta-S2", "one-cta-comb2", "one-cta-comb2-S2", "slim-type2", "one-cta-S4",
                              "unstaged-dense-S2", "unstaged-341prb-S4"])
def test_persistent_loop_full_band_layouts(pattern, syms, n_units, grid):
    """Full-band geometries that select the other shared-memory layouts (single staging buffer, per-group FIR
    input, one 704-thread CTA per SM), with more units than resident CTAs so that every CTA pipelines several."""
    from srsran_ce_pytorch_b200 import estimator, synth
    hop1 = synth.make_hop(grid, 0, grid, syms, pattern, 2)
    cfg = synth.make_config(30e3, "filter", True)
    gen = 20
    data = synth.make_units(hop1, synth.no_hop(), cfg, 4, gen, pilot_share=4, seed=21, device="cuda")
    reps = n_units // gen
    rx, pil = data["rx"].repeat(reps, 1, 1), data["pilots"].repeat(reps, 1, 1, 1)
    out, sc = estimator.srs_channel_estimator_batched(rx, pil, 1.0, hop1, synth.no_hop(), cfg, pilot_share=4)
    # every repetition of the 20 generated units must be bit-identical, wherever in the launch it was processed
    ref_out, ref_sc = out[:gen], sc[:gen]
    for r in (1, reps // 2, reps - 1):
        assert torch.equal(out[r * gen:(r + 1) * gen], ref_out) and torch.equal(sc[r * gen:(r + 1) * gen], ref_sc), r
    rx_h, pil_h = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
    for b in (0, 7, 19):
        diag = {}
        want = ce_oracle.estimate(rx_h[b], pil_h[b // 4], 1.0, hop1, synth.no_hop(), cfg, diagnostics=diag)
        mx, nmse = tol.grid_errors(out[(reps - 1) * gen + b].cpu().numpy(), want[0])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (pattern, syms, b, mx, nmse)
        s = sc[(reps - 1) * gen + b].cpu().numpy()
        tol.check_scalars([s[0], s[1], s[2], s[3], s[4] if s[5] else None], want[1:], 30e3, diag["ta_margin"],
                          tag=f"layouts/{pattern}/{b}")


def test_indexed_received_grids_equal_materialised_copies():
    """srsce_estimate_indexed: unit u reads rx_grids[rx_index[u]] -- bit-identical to gathering the grids first."""
    from srsran_ce_pytorch_b200 import estimator
    hop1, hop2, cfg, data = _config2(32)
    plan = estimator.make_plan(273, 14, 4, hop1, hop2, cfg)
    grids = data["rx"][:4].contiguous()                          # 4 rx antennas of one slot ...
    idx = (torch.arange(32, device="cuda") % 4).to(torch.int64)  # ... shared by the 8 UEs (unit = ue * 4 + antenna)
    a, sa = plan.estimate_indexed(grids, idx, data["pilots"], 1.0, 4)
    b, sb = plan.estimate(grids[idx].contiguous(), data["pilots"], 1.0, 4)
    # The grids are bit-identical.  The scalars are float32 partial sums per estimator thread combined in float64, and
    # the shared-grid launch runs the plan's CTA-pair layout (256 estimator threads) where the per-unit-grid launch runs
    # 512: the summation order differs, so they agree to float32 rounding, not bitwise.
    assert torch.equal(a, b)
    assert torch.allclose(sa, sb, rtol=2e-6, atol=1e-12)
    with pytest.raises(IndexError):
        plan.estimate_indexed(grids, idx + 1, data["pilots"], 1.0, 4)


def test_host_pipeline_matches_device_path():
    """hostio.HostPipeline (pinned host buffers in / out, two streams) returns what Plan.estimate computes, for the
    dense and the compact output, including a last partial chunk."""
    from srsran_ce_pytorch_b200 import estimator
    from srsran_ce_pytorch_b200.hostio import HostPipeline
    hop1, hop2, cfg, data = _config2(40, syms=(3, 10))
    plan = estimator.make_plan(273, 14, 4, hop1, hop2, cfg)
    dense, sc = plan.estimate(data["rx"], data["pilots"], 1.0, 4)
    h_rx, h_pil = data["rx"].cpu().pin_memory(), data["pilots"].cpu().pin_memory()
    h_out = torch.empty(dense.shape, dtype=torch.complex64).pin_memory()
    h_sc = torch.empty(sc.shape, dtype=torch.float64).pin_memory()
    pipe = HostPipeline(plan, 16, 4)
    pipe.run(h_rx, h_pil, 1.0, h_out, h_sc)
    assert torch.equal(h_out, dense.cpu()) and torch.equal(h_sc, sc.cpu())
    assert pipe.launches == 3 and pipe.d2h_bytes == dense.numel() * 8 + 40 * 48
    ch, rot, _ = plan.estimate_compact(data["rx"], data["pilots"], 1.0, 4)
    hc = (torch.empty(ch.shape, dtype=torch.complex64).pin_memory(), torch.empty(rot.shape, dtype=torch.complex64).pin_memory())
    HostPipeline(plan, 16, 4, compact=True).run(h_rx, h_pil, 1.0, hc, h_sc)
    assert torch.equal(hc[0], ch.cpu()) and torch.equal(hc[1], rot.cpu()) and torch.equal(h_sc, sc.cpu())


@pytest.mark.parametrize("two_hops", [False, True], ids=["1hop", "2hop"])
def test_host_pipeline_with_host_side_replication(two_hops):
    """HostPipeline(replicate="host"): the per-hop columns cross PCIe and srsce_replicate_host rebuilds the dense grid
    in the caller's host buffer -- bit-identical to the grid the fused kernel writes on the device (the host side only
    copies); the scalars agree to float32 summation order (the compact launch may use another CTA shape).  Plans with
    a final CFO rotation are refused ("auto" falls back to the dense device->host copy for them)."""
    from srsran_ce_pytorch_b200 import estimator, synth
    from srsran_ce_pytorch_b200.hostio import HostPipeline, srs_channel_estimator_host
    if two_hops:
        hop1 = synth.make_hop(52, 2, 24, [2], "comb2", 2, start_symbol=0, n_alloc=7)
        hop2 = synth.make_hop(52, 20, 24, [9], "comb2", 2, start_symbol=7, n_alloc=7)
        cfg = synth.make_config(30e3, "filter", True)
        data = synth.make_units(hop1, hop2, cfg, 4, 40, pilot_share=4, seed=3, device="cuda")
        plan = estimator.make_plan(52, 14, 4, hop1, hop2, cfg)
    else:
        hop1, hop2, cfg, data = _config2(40, syms=(13,))
        plan = estimator.make_plan(273, 14, 4, hop1, hop2, cfg)
    assert not plan.has_final_rotation
    dense, sc = plan.estimate(data["rx"], data["pilots"], 1.0, 4)
    h_rx, h_pil = data["rx"].cpu().pin_memory(), data["pilots"].cpu().pin_memory()
    h_out = torch.full(dense.shape, float("nan"), dtype=torch.complex64).pin_memory()
    h_sc = torch.empty(sc.shape, dtype=torch.float64).pin_memory()
    for streams in (2, 3):
        pipe = HostPipeline(plan, 16, 4, n_streams=streams, replicate="host")
        h_out.fill_(float("nan"))
        pipe.run(h_rx, h_pil, 1.0, h_out, h_sc)
        assert torch.equal(torch.view_as_real(h_out), torch.view_as_real(dense.cpu()))
        assert torch.allclose(h_sc, sc.cpu(), rtol=2e-6, atol=1e-12)
        assert torch.equal(h_sc[:, 3], sc.cpu()[:, 3])                          # time alignment: exact index
        assert pipe.launches == 3 and pipe.d2h_bytes == 40 * plan.info.n_sc * 4 * 8 * (2 if two_hops else 1) + 40 * 48
    out2, sc2 = srs_channel_estimator_host(plan, h_rx, h_pil, 1.0, 4, chunk_units=16)      # "auto" picks the host side here
    assert torch.equal(torch.view_as_real(out2), torch.view_as_real(dense.cpu())) and torch.allclose(sc2, sc.cpu(), rtol=2e-6, atol=1e-12)
    # a plan with the final rotation: host replication refused, "auto" takes the dense copy
    hop1r, hop2r, cfgr, datar = _config2(8, syms=(3, 10))
    planr = estimator.make_plan(273, 14, 4, hop1r, hop2r, cfgr)
    assert planr.has_final_rotation
    with pytest.raises(NotImplementedError):
        HostPipeline(planr, 8, 4, replicate="host")
    assert HostPipeline(planr, 8, 4, replicate="auto").replicate == "device"


@pytest.mark.parametrize("global_h", [False, True], ids=["smem-h", "global-h"])
def test_launch_is_capturable_in_a_cuda_graph(global_h):
    """The library never synchronises and touches only the stream it is given, so a call can be captured in a CUDA
    graph (launch-bound small batches) and replayed on new data in the same buffers -- also for the layout-F plan, whose
    per-launch scratch is a stream-ordered allocation (a memory node of the graph)."""
    from srsran_ce_pytorch_b200 import estimator, synth
    if global_h:
        hop1 = synth.make_hop(273, 0, 273, [3], "dense", 2, start_symbol=0, n_alloc=7)
        hop2 = synth.make_hop(273, 0, 273, [10], "dense", 2, start_symbol=7, n_alloc=7)
        grid = 273
    else:
        hop1, hop2, grid = synth.make_hop(52, 40, 3, [0, 4, 8, 12], "comb2", 1), synth.no_hop(), 52   # the config-1 shape
    L = 4 if global_h else 1
    cfg = synth.make_config(30e3 if global_h else 15e3, "filter", True)
    plan = estimator.make_plan(grid, 14, L, hop1, hop2, cfg)
    a = synth.make_units(hop1, hop2, cfg, L, 8, seed=1, device="cuda")
    b = synth.make_units(hop1, hop2, cfg, L, 8, seed=2, device="cuda")
    want_a, sc_a = plan.estimate(a["rx"], a["pilots"], 1.0, 1)
    want_b, sc_b = plan.estimate(b["rx"], b["pilots"], 1.0, 1)
    rx, pil = a["rx"].clone(), a["pilots"].clone()
    out, sc = torch.empty_like(want_a), torch.empty_like(sc_a)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        plan.estimate(rx, pil, 1.0, 1, out=out, scalars=sc)              # warm-up outside the capture
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan.estimate(rx, pil, 1.0, 1, out=out, scalars=sc)
    out.zero_()
    g.replay()
    assert torch.equal(out, want_a) and torch.equal(sc, sc_a)
    rx.copy_(b["rx"])
    pil.copy_(b["pilots"])
    g.replay()
    assert torch.equal(out, want_b) and torch.equal(sc, sc_b)


@pytest.mark.parametrize("global_h", [False, True], ids=["smem-h", "global-h"])
def test_one_plan_shared_by_threads_and_streams(global_h):
    """include/srsce.h: a plan is immutable and may be shared by threads and streams.  Four threads, each on its own
    stream, run the same plan on different inputs at the same time (for the layout-F plan every launch takes its own
    stream-ordered scratch); every result equals the one computed alone."""
    import threading
    from srsran_ce_pytorch_b200 import estimator, synth
    if global_h:
        hop1 = synth.make_hop(273, 0, 273, [3], "dense", 2, start_symbol=0, n_alloc=7)
        hop2 = synth.make_hop(273, 0, 273, [10], "dense", 2, start_symbol=7, n_alloc=7)
        n = 160
    else:
        hop1, hop2, n = synth.make_hop(273, 0, 273, [13], "comb4", 2), synth.no_hop(), 320
    cfg = synth.make_config(30e3, "filter", True)
    plan = estimator.make_plan(273, 14, 4, hop1, hop2, cfg)
    assert bool(plan.info.h_in_global) == global_h
    inputs = [synth.make_units(hop1, hop2, cfg, 4, n, pilot_share=4, seed=20 + k, device="cuda") for k in range(4)]
    alone = [plan.estimate(d["rx"], d["pilots"], 1.0, 4) for d in inputs]
    torch.cuda.synchronize()
    results, errors = [None] * 4, []

    def work(k):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(3):
                    results[k] = plan.estimate(inputs[k]["rx"], inputs[k]["pilots"], 1.0, 4)
            st.synchronize()
        except Exception as e:          # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(4):
        assert torch.equal(results[k][0], alone[k][0]) and torch.equal(results[k][1], alone[k][1]), k


@pytest.mark.parametrize("syms", [(13,), (3, 10)], ids=["single-cta", "cta-pair"])
def test_unit_counts_around_the_resident_cta_count(syms):
    """Launches of 1 ... 2 x SMs + 1 units: grids smaller than, equal to and just above the number of resident CTAs
    (one per SM for the single-CTA layout, two for the pair) give the same per-unit results as one large launch."""
    from srsran_ce_pytorch_b200 import estimator
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    hop1, hop2, cfg, data = _config2(2 * n_sm + 4, syms=syms)
    plan = estimator.make_plan(273, 14, 4, hop1, hop2, cfg)
    full, sc_full = plan.estimate(data["rx"], data["pilots"], 1.0, 4)
    for n in (4, n_sm - 4, n_sm, n_sm + 4, 2 * n_sm - 4, 2 * n_sm, 2 * n_sm + 4):
        out, sc = plan.estimate(data["rx"][:n], data["pilots"][:n // 4], 1.0, 4)
        assert torch.equal(out, full[:n]) and torch.equal(sc, sc_full[:n]), n


def test_lazy_conjugate_and_negative_views_are_resolved():
    """torch keeps ``x.conj()`` as a flag on the same storage: the kernel reads raw storage, so the host layer must
    materialise such views.  Results must equal the oracle on the conjugated values (as the reference computes)."""
    from srsran_ce_pytorch_b200 import estimator
    hop1, hop2, cfg, data = _config2(4, syms=(3, 10))
    rx, pil = data["rx"], data["pilots"]
    rx_c, pil_c = rx.conj(), pil.conj()
    assert rx_c.is_conj() and rx_c.data_ptr() == rx.data_ptr()
    out, sc = estimator.srs_channel_estimator_batched(rx_c, pil_c, 1.0, hop1, hop2, cfg, pilot_share=4)
    want = ce_oracle.estimate(rx[1].conj().resolve_conj().cpu().numpy(), pil[0].conj().resolve_conj().cpu().numpy(), 1.0,
                              hop1, hop2, cfg)
    mx, nmse = tol.grid_errors(out[1].cpu().numpy(), want[0])
    assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (mx, nmse)
    # a double conjugate and a negated view are the identity / the negated grid
    ref, _ = estimator.srs_channel_estimator_batched(rx, pil, 1.0, hop1, hop2, cfg, pilot_share=4)
    twice, _ = estimator.srs_channel_estimator_batched(rx.conj().conj(), pil.conj().conj(), 1.0, hop1, hop2, cfg, pilot_share=4)
    assert torch.equal(ref, twice)
    neg = torch.view_as_complex(torch._neg_view(torch.view_as_real(rx)))
    assert neg.is_neg()
    out_n, _ = estimator.srs_channel_estimator_batched(neg, pil, 1.0, hop1, hop2, cfg, pilot_share=4)
    mx, nmse = tol.grid_errors(out_n.cpu().numpy(), (-ref).cpu().numpy())   # (virtual pilots: phase + pi, not bit exact)
    assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (mx, nmse)
    one = estimator.srs_channel_estimator(rx_c[1], pil_c[0], 1.0, hop1, hop2, cfg)
    assert torch.equal(one[0], out[1])


def test_caller_supplied_result_buffers_are_validated():
    """A wrongly shaped / typed / strided ``out`` or ``scalars`` would be overrun by the kernel: every compute call refuses it."""
    from srsran_ce_pytorch_b200 import estimator
    hop1, hop2, cfg, data = _config2(8)
    plan = estimator.make_plan(273, 14, 4, hop1, hop2, cfg)
    idx = torch.arange(8, device="cuda", dtype=torch.int64)
    good_out = torch.empty((8, 3276, 14, 4), dtype=torch.complex64, device="cuda")
    good_sc = torch.empty((8, 6), dtype=torch.float64, device="cuda")
    bad_outs = [good_out[:4], torch.empty((8, 3276, 14, 4), dtype=torch.complex128, device="cuda"),
                torch.empty((8, 3276, 4, 14), dtype=torch.complex64, device="cuda").transpose(2, 3),
                torch.empty((8, 3276, 14, 4), dtype=torch.complex64)]
    for bad in bad_outs:
        with pytest.raises(ValueError):
            plan.estimate(data["rx"], data["pilots"], 1.0, 4, out=bad)
        with pytest.raises(ValueError):
            plan.estimate_indexed(data["rx"], idx, data["pilots"], 1.0, 4, out=bad)
    for bad in (good_sc[:, :5], good_sc.float(), good_sc.t().contiguous().t()):
        with pytest.raises(ValueError):
            plan.estimate(data["rx"], data["pilots"], 1.0, 4, scalars=bad)
        with pytest.raises(ValueError):
            plan.estimate_indexed(data["rx"], idx, data["pilots"], 1.0, 4, scalars=bad)
    a, sa = plan.estimate_indexed(data["rx"], idx, data["pilots"], 1.0, 4, out=good_out, scalars=good_sc, check_index=False)
    assert a.data_ptr() == good_out.data_ptr() and sa.data_ptr() == good_sc.data_ptr()


def test_plan_memo_skips_host_reads_and_sees_in_place_changes():
    """Repeated drop-in calls with the same HopConfig / EstimatorConfig objects reuse the plan without touching their
    tensors; an in-place change of a mask tensor (version counter) is noticed."""
    from srsran_ce_pytorch_b200 import estimator, synth
    hop1 = synth.make_hop(52, 2, 8, [1, 3], "comb4", 2)
    hop2, cfg = synth.no_hop(), synth.make_config(30e3, "filter", True)
    hop1.DMRSsymbols = hop1.DMRSsymbols.cuda()
    hop1.maskPRBs = hop1.maskPRBs.cuda()
    p1 = estimator.make_plan(52, 14, 4, hop1, hop2, cfg)
    assert estimator.make_plan(52, 14, 4, hop1, hop2, cfg) is p1
    hop1.DMRSsymbols[5] = True
    p2 = estimator.make_plan(52, 14, 4, hop1, hop2, cfg)
    assert p2 is not p1 and p2.info.n_dmrs_total == 3
