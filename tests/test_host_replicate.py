"""srsce_replicate_host (the host half of the host-buffer pipeline): the per-hop columns of the compact output,
replicated over the allocated symbols by the library's host threads, equal ``Plan.expand_compact`` (plain torch ops)
bit for bit -- it is a copy.  Runs without a GPU: the call needs no device."""
import pytest
import torch

from srsran_ce_pytorch_b200 import estimator, synth

CASES = [
    # grid, hop1 (start, n, dmrs syms, start_symbol, n_alloc), hop2 or None, layers, n_sym
    (52, (3, 40, [2], 0, 14), None, 4, 14),              # the common shape: straight streaming path, 32-byte rows
    (52, (0, 52, [2], 0, 14), None, 2, 14),              # 16-byte rows
    (24, (2, 20, [3], 0, 14), None, 1, 14),              # 8-byte rows: block path
    (24, (2, 20, [3], 0, 14), None, 3, 14),              # 24-byte rows
    (24, (0, 24, [3], 1, 12), None, 4, 14),              # unallocated symbols are zero
    (24, (0, 24, [6], 0, 13), None, 3, 13),              # odd line size: unaligned lines
    (24, (0, 8, [2], 0, 7), (12, 8, [9], 7, 7), 2, 14),  # two hops, disjoint bands
    (24, (4, 12, [2], 1, 5), (8, 12, [9], 6, 7), 4, 14), # two hops, overlapping bands, partial allocation
    (24, (4, 12, [2], 1, 5), (8, 12, [9], 6, 7), 1, 14),
]


def _plan(case):
    grid, h1, h2, L, n_sym = case
    n_cdm = (L + 1) // 2
    hop1 = synth.make_hop(grid, h1[0], h1[1], h1[2], "comb2", n_cdm, n_sym=n_sym, start_symbol=h1[3], n_alloc=h1[4])
    hop2 = synth.no_hop() if h2 is None else synth.make_hop(grid, h2[0], h2[1], h2[2], "comb2", n_cdm, n_sym=n_sym,
                                                            start_symbol=h2[3], n_alloc=h2[4])
    cfg = synth.make_config(30e3, "filter", False)
    return estimator.make_plan(grid, n_sym, L, hop1, hop2, cfg, "linear")


def _columns(plan, n_units, seed):
    """Random per-hop columns that are zero outside each hop's band (what srsce_estimate_compact returns)."""
    i, d = plan.info, plan.desc
    n_hops = 2 if d.hop[1].dmrs_symbol_mask else 1
    g = torch.Generator().manual_seed(seed)
    ch = torch.view_as_complex(torch.randn((n_units, n_hops, i.n_sc, i.n_layers, 2), generator=g))
    for hh in range(n_hops):
        band = torch.zeros(i.n_sc, dtype=torch.bool)
        band[12 * d.hop[hh].prb_start:12 * (d.hop[hh].prb_start + d.hop[hh].n_prbs)] = True
        ch[:, hh, ~band, :] = 0
    return ch.contiguous()


@pytest.mark.parametrize("case", CASES, ids=[f"c{k}" for k in range(len(CASES))])
@pytest.mark.parametrize("n_units,threads", [(1, 1), (3, 0), (37, 4)])
def test_replicate_host_equals_expand_compact(case, n_units, threads):
    plan = _plan(case)
    assert not plan.has_final_rotation
    i = plan.info
    ch = _columns(plan, n_units, seed=n_units)
    want = plan.expand_compact(ch, torch.ones((n_units, i.n_sym), dtype=torch.complex64))
    out = torch.full((n_units, i.n_sc, i.n_sym, i.n_layers), float("nan"), dtype=torch.complex64)
    got = plan.replicate_host(ch, out, threads)
    assert got is out
    assert torch.equal(torch.view_as_real(out), torch.view_as_real(want))


def test_replicate_host_unaligned_destination_and_repeated_calls():
    """The grid may start on any 8-byte boundary (a slice of a larger host array); the pool is reused across calls."""
    plan = _plan(CASES[0])
    i = plan.info
    ch = _columns(plan, 5, seed=1)
    want = plan.expand_compact(ch, torch.ones((5, i.n_sym), dtype=torch.complex64))
    flat = torch.zeros(5 * i.n_sc * i.n_sym * i.n_layers + 1, dtype=torch.complex64)
    out = flat[1:].view(5, i.n_sc, i.n_sym, i.n_layers)
    assert out.data_ptr() % 16 == 8 or out.data_ptr() % 16 == 0
    for _ in range(3):
        out.zero_()
        plan.replicate_host(ch, out, 0)
        assert torch.equal(torch.view_as_real(out), torch.view_as_real(want))
    assert flat[0] == 0


def test_replicate_host_refuses_plans_with_a_final_rotation_and_bad_buffers():
    hop1 = synth.make_hop(24, 0, 24, [2, 11], "comb2", 1)
    plan = estimator.make_plan(24, 14, 2, hop1, synth.no_hop(), synth.make_config(30e3, "filter", True), "linear")
    assert plan.has_final_rotation
    i = plan.info
    ch = torch.zeros((1, 1, i.n_sc, 2), dtype=torch.complex64)
    out = torch.zeros((1, i.n_sc, 14, 2), dtype=torch.complex64)
    with pytest.raises(NotImplementedError):
        plan.replicate_host(ch, out)
    ch = _columns(plan, 1, seed=2)                                         # columns of srsce_fill_compact: never rotated
    plan.replicate_host(ch, out, from_fill=True)
    assert torch.equal(torch.view_as_real(out), torch.view_as_real(plan.expand_compact(ch, torch.ones((1, 14), dtype=torch.complex64))))
    plan2 = _plan(CASES[1])
    i2 = plan2.info
    with pytest.raises(IndexError):
        plan2.replicate_host(torch.zeros((1, 1, i2.n_sc, 3), dtype=torch.complex64),
                             torch.zeros((1, i2.n_sc, 14, 2), dtype=torch.complex64))
    with pytest.raises(ValueError):
        plan2.replicate_host(torch.zeros((1, 1, i2.n_sc, 2), dtype=torch.complex128),
                             torch.zeros((1, i2.n_sc, 14, 2), dtype=torch.complex64))
    empty = plan2.replicate_host(torch.zeros((0, 1, i2.n_sc, 2), dtype=torch.complex64),
                                 torch.zeros((0, i2.n_sc, 14, 2), dtype=torch.complex64))
    assert empty.shape[0] == 0
