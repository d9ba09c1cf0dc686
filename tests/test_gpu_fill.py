"""srsce_fill (the fill_ch_est_cdm seam alone): the direct and the tcgen05 tensor-core implementation against the
oracle's band fill on seeded inputs, against outputs of the unmodified reference (tests/golden_fill) through the
drop-in ``fill_ch_est_cdm``, and size-independent properties at BASELINE config 4's full size.
Floating point: max relative error 1e-5 and NMSE 1e-12 like the estimator (tests/tolerances.py); the two
implementations agree with each other to 2e-6."""
import numpy as np
import pytest
import torch

from oracle import ce_oracle
from tests import tolerances as tol
from tests.test_oracle_fill_golden import NAMES as FILL_NAMES, load_fill_case

pytestmark = pytest.mark.gpu


def _expected(h, hop1, hop2, n_sc, n_sym, fill):
    """h [B, n_hops, N, L] numpy -> dense grids via the oracle's band fill (hop 2 wins where the hops overlap)."""
    B, n_hops, N, L = h.shape
    out = np.zeros((B, n_sc, n_sym, L), dtype=np.complex64)
    for hh, hop in enumerate((hop1, hop2)[:n_hops]):
        spec = ce_oracle.hop_from_any(hop)
        sc0 = 12 * spec.prb_start
        for b in range(B):
            for g in range((L + 1) // 2):
                band = ce_oracle._fill_band(h[b, hh][:, 2 * g:2 * g + 2], spec, g, "cnn" if fill == "cnn" else "tensorized")
                for l in range(band.shape[1]):
                    out[b, sc0:sc0 + band.shape[0], spec.start_symbol:spec.start_symbol + spec.n_alloc_symbols, 2 * g + l] = band[:, l:l + 1]
    return out


FILL_CASES = [
    # grid PRBs, start, PRBs, pattern, layers, start symbol, allocated symbols, units
    (52, 0, 52, "comb4", 4, 0, 14, 37),
    (52, 5, 24, "comb2", 4, 2, 9, 20),
    (52, 3, 17, "comb4", 2, 0, 14, 70),
    (52, 44, 8, "comb2", 1, 0, 14, 33),
    (52, 10, 9, "comb4", 3, 1, 12, 34),
    (52, 0, 30, "comb6", 4, 0, 14, 40),
    (52, 7, 16, "comb12", 2, 0, 14, 5),
    (25, 2, 11, "dense", 2, 0, 14, 19),
    (52, 20, 1, "comb4", 4, 0, 14, 3),
    (52, 9, 17, "type2", 1, 0, 14, 18),
    (106, 1, 100, "comb3", 2, 3, 8, 17),
]


@pytest.mark.parametrize("case", FILL_CASES, ids=[f"f{i}" for i in range(len(FILL_CASES))])
@pytest.mark.parametrize("fill", ["linear", "cnn"])
@pytest.mark.parametrize("impl", ["direct", "tensor"])
def test_fill_matches_oracle(case, fill, impl):
    from srsran_ce_pytorch_b200 import estimator, synth
    grid, start, n, pat, L, s0, n_alloc, B = case
    hop1 = synth.make_hop(grid, start, n, [2], pat, (L + 1) // 2, 14, s0, n_alloc)
    plan = estimator.make_plan(grid, 14, L, hop1, synth.no_hop(), synth.make_config(30e3, "filter", True), fill)
    if impl == "tensor" and not plan.info.fill_tensor:
        pytest.skip("no tensor-core fill for this geometry")
    g = torch.Generator().manual_seed(17)
    h = torch.complex(torch.randn(B, 1, plan.info.n_re, L, generator=g), torch.randn(B, 1, plan.info.n_re, L, generator=g)).to(torch.complex64)
    out = plan.fill(h.cuda(), impl=impl).cpu().numpy()
    want = _expected(h.numpy(), hop1, None, 12 * grid, 14, fill)
    for b in range(B):
        mx, nmse = tol.grid_errors(out[b], want[b])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (case, fill, impl, b, mx, nmse)
    # exact zeros outside the band and the allocated symbols
    mask = np.ones((12 * grid, 14), dtype=bool)
    mask[12 * start:12 * (start + n), s0:s0 + n_alloc] = False
    assert not out[:, mask].any()


@pytest.mark.parametrize("fill", ["linear", "cnn"])
def test_fill_direct_two_hops(fill):
    from srsran_ce_pytorch_b200 import estimator, synth
    hop1 = synth.make_hop(52, 4, 12, [2], "comb4", 2, 14, 0, 7)
    hop2 = synth.make_hop(52, 10, 12, [9], "comb4", 2, 14, 5, 9)        # overlapping PRBs and symbols: hop 2 wins
    plan = estimator.make_plan(52, 14, 4, hop1, hop2, synth.make_config(30e3, "filter", True), fill)
    g = torch.Generator().manual_seed(3)
    h = torch.complex(torch.randn(9, 2, plan.info.n_re, 4, generator=g), torch.randn(9, 2, plan.info.n_re, 4, generator=g)).to(torch.complex64)
    out = plan.fill(h.cuda()).cpu().numpy()
    want = _expected(h.numpy(), hop1, hop2, 624, 14, fill)
    for b in range(9):
        mx, nmse = tol.grid_errors(out[b], want[b])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (fill, b, mx, nmse)
    with pytest.raises(NotImplementedError):
        plan.fill(h.cuda(), impl="tensor")


@pytest.mark.parametrize("name", FILL_NAMES)
@pytest.mark.parametrize("impl", ["direct", "tensor"])
def test_drop_in_fill_ch_est_cdm_matches_reference(name, impl):
    """estimator.fill_ch_est_cdm(channel_in, estimated, hop, i_cdm) vs the stored outputs of the reference's
    ce_rule_tensorized / ce_rule_baseline / ce_dl_cnn functions of the same name."""
    from srsran_ce_pytorch_b200 import estimator
    from srsran_ce_pytorch_b200.types import HopConfig
    z, spec = load_fill_case(name)
    hop = HopConfig(DMRSsymbols=torch.from_numpy(z["dmrs_symbols"]), DMRSREmask=torch.from_numpy(z["re_mask"]),
                    PRBstart=spec.prb_start, nPRBs=spec.n_prbs, maskPRBs=torch.from_numpy(z["mask_prbs"]),
                    startSymbol=spec.start_symbol, nAllocatedSymbols=spec.n_alloc_symbols)
    ch, est = torch.from_numpy(z["channel_in"]).cuda(), torch.from_numpy(z["estimated"]).cuda()
    for fill, refs in (("linear", ("tensorized", "baseline")), ("cnn", ("cnn",))):
        try:
            out = estimator.fill_ch_est_cdm(ch, est, hop, int(z["i_cdm"]), fill=fill, impl=impl)
        except NotImplementedError:
            assert impl == "tensor"
            continue
        assert out.data_ptr() != ch.data_ptr() and out.shape == ch.shape and out.dtype == torch.complex64
        for r in refs:
            mx, nmse = tol.grid_errors(out.cpu().numpy(), z[f"out_{r}"])
            assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (name, impl, fill, r, mx, nmse)


@pytest.mark.parametrize("fill", ["cnn", "linear"])
def test_config4_full_size_tensor_equals_direct(fill):
    """BASELINE config 4 shape (273 PRB, comb-4, 4 ports; 300 UE-antenna grids here): the tcgen05 path and the direct
    path agree to float32 rounding, the fill is linear, and known pilots are reproduced exactly by the CNN fill."""
    from srsran_ce_pytorch_b200 import estimator, synth
    hop1 = synth.make_hop(273, 0, 273, [13], "comb4", 2)
    plan = estimator.make_plan(273, 14, 4, hop1, synth.no_hop(), synth.make_config(30e3, "filter", True), fill)
    assert plan.info.fill_tensor == 1
    g = torch.Generator().manual_seed(5)
    h = torch.complex(torch.randn(300, 1, 819, 4, generator=g), torch.randn(300, 1, 819, 4, generator=g)).to(torch.complex64).cuda()
    a, b = plan.fill(h, impl="direct"), plan.fill(h, impl="tensor")
    assert a.shape == (300, 3276, 14, 4)
    scale = float(a.abs().max())
    assert float((a - b).abs().max()) <= 2e-6 * scale
    assert torch.equal(b[:, :, 0, :], b[:, :, 13, :])                    # replicated over the allocated symbols
    # linearity: fill(2 h1 - h2) == 2 fill(h1) - fill(h2)
    c = plan.fill(2 * h[:100] - h[100:200], impl="tensor")
    assert float((c - (2 * b[:100] - b[100:200])).abs().max()) <= 4e-6 * scale
    if fill == "cnn":                                                    # known samples are re-imposed (ce_dl_cnn.py:505-508)
        pos0 = torch.arange(0, 3276, 4, device="cuda")                  # CDM group 0 of comb-4: rows 0, 4, 8, ...
        assert torch.equal(a[:, pos0, 0, 0], h[:, 0, :, 0]) and torch.equal(a[:, pos0, 0, 1], h[:, 0, :, 1])
        # (tensor path: 1.0 * (x0 + x1 + x2) accumulated in the tensor core's float32 adder -- equal to the last bit or so)
        assert float((b[:, pos0, 0, 0] - h[:, 0, :, 0]).abs().max()) <= 2.5e-7 * scale
    want = _expected(h[:2].cpu().numpy(), hop1, None, 3276, 14, fill)
    for bb in range(2):
        mx, nmse = tol.grid_errors(b[bb].cpu().numpy(), want[bb])
        assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (bb, mx, nmse)


def _fuzz_cases(n, seed):
    import random
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        r = rnd.choice([1, 2, 3, 4, 5, 6, 7, 12])
        layers = rnd.choice([1, 2, 3, 4])
        cols = []
        for _ in range((layers + 1) // 2):
            pos = sorted(rnd.sample(range(12), r))
            cols.append("".join("1" if k in pos else "0" for k in range(12)))
        grid = rnd.choice([25, 52, 106, 273])
        n_prb = min(rnd.choice([1, 2, 3, 7, 8, 9, 16, 17, 40, 100, 273]), grid)
        start = rnd.randrange(0, grid - n_prb + 1)
        s0 = rnd.choice([0, 0, 3])
        n_alloc = rnd.choice([14 - s0, 14 - s0, max(1, 9 - s0)])
        out.append((grid, start, n_prb, cols, layers, s0, n_alloc, rnd.choice([1, 7, 33, 65])))
    return out


FUZZ = _fuzz_cases(40, 4242)


@pytest.mark.parametrize("k", range(len(FUZZ)))
def test_fill_random_masks_match_oracle(k):
    """Arbitrary 12-bit RE masks (irregular, R not dividing 12), band positions and widths that end tiles anywhere, odd
    layer counts, partial allocations, unit counts that leave partial MMA batches: both implementations vs the oracle."""
    from srsran_ce_pytorch_b200 import estimator, synth
    grid, start, n, cols, L, s0, n_alloc, B = FUZZ[k]
    hop1 = synth.make_hop(grid, start, n, [2], cols, (L + 1) // 2, 14, s0, n_alloc)
    for fill in ("linear", "cnn"):
        try:
            plan = estimator.make_plan(grid, 14, L, hop1, synth.no_hop(), synth.make_config(30e3, "filter", True), fill)
        except NotImplementedError:
            continue                                          # (CNN operator rows wider than the kernel accepts: some irregular masks)
        g = torch.Generator().manual_seed(100 + k)
        h = torch.complex(torch.randn(B, 1, plan.info.n_re, L, generator=g), torch.randn(B, 1, plan.info.n_re, L, generator=g)).to(torch.complex64)
        want = _expected(h[:2].numpy(), hop1, None, 12 * grid, 14, fill)
        for impl in ("direct", "tensor"):
            if impl == "tensor" and not plan.info.fill_tensor:
                continue
            out = plan.fill(h.cuda(), impl=impl)
            for b in range(min(2, B)):
                mx, nmse = tol.grid_errors(out[b].cpu().numpy(), want[b])
                assert mx <= tol.GRID_MAX_REL and nmse <= tol.GRID_NMSE, (FUZZ[k], fill, impl, b, mx, nmse)
            if B > 2:                                         # batch independence: the last unit equals a batch of its own
                solo = plan.fill(h[-1:].cuda(), impl=impl)
                assert torch.equal(solo[0], out[-1]), (FUZZ[k], fill, impl)


@pytest.mark.parametrize("case", [FILL_CASES[0], FILL_CASES[1], FILL_CASES[4], FILL_CASES[9]], ids=["f0", "f1", "f4", "f9"])
@pytest.mark.parametrize("fill", ["linear", "cnn"])
def test_fill_compact_and_host_pipeline(case, fill):
    """srsce_fill_compact leaves the fill as per-hop columns; replicated over the allocated symbols (on the device by
    ``expand_compact``, on the host by ``srsce_replicate_host``) they are bit-identical to srsce_fill's dense grid.
    FillHostPipeline (host arrays in / out) returns the same grid in both replication modes, including a partial
    last chunk, and for a plan whose estimator output would carry a CFO rotation (the fill never rotates)."""
    from srsran_ce_pytorch_b200 import estimator, synth
    from srsran_ce_pytorch_b200.hostio import FillHostPipeline
    grid, start, n, pat, L, s0, na, units = case
    hop1 = synth.make_hop(grid, start, n, [s0 + 1, s0 + 3] if na >= 4 else [s0], pat, (L + 1) // 2, start_symbol=s0, n_alloc=na)
    cfg = synth.make_config(30e3, "filter", True)
    plan = estimator.make_plan(grid, 14, L, hop1, synth.no_hop(), cfg, fill)
    i = plan.info
    g = torch.Generator(device="cuda").manual_seed(3)
    h = torch.view_as_complex(torch.randn((units, 1, i.n_re, L, 2), device="cuda", generator=g))
    dense = plan.fill(h, impl="direct")
    ch = plan.fill_compact(h)
    assert ch.shape == (units, 1, i.n_sc, L)
    ones = torch.ones((units, 14), dtype=torch.complex64, device="cuda")
    assert torch.equal(torch.view_as_real(plan.expand_compact(ch, ones)), torch.view_as_real(dense))
    h_host = h.cpu().pin_memory()
    out = torch.full(dense.shape, float("nan"), dtype=torch.complex64).pin_memory()
    for mode in ("host", "device"):
        pipe = FillHostPipeline(plan, 16, n_streams=2, replicate=mode)
        out.fill_(float("nan"))
        pipe.run(h_host, out)
        assert torch.equal(torch.view_as_real(out), torch.view_as_real(dense.cpu())), mode
        assert pipe.launches == (units + 15) // 16
    if plan.has_final_rotation:
        with pytest.raises(NotImplementedError):
            plan.replicate_host(ch.cpu(), out)                    # estimator columns of this plan would need the rotation
