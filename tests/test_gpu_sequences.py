"""srsce_gen_pilots (an extension: the reference has no sequence generator, parity unpinned).  Checked
against an independent float64 restatement of TS 38.211 5.2.2 written here, by properties of the sequences, and end to
end: generated pilots through a known channel come back from the estimator."""
import ctypes as C

import numpy as np
import pytest
import torch

from srsran_ce_pytorch_b200 import _lib


def _spec_sequence(u, v, m):
    """r_{u,v}(n), 0 <= n < m, straight from the formulas of TS 38.211 5.2.2.1 / 5.2.2.2 in float64."""
    n = np.arange(m, dtype=np.float64)
    if m == 30:
        return np.exp(-1j * np.pi * (u + 1) * (n + 1) * (n + 2) / 31.0)
    n_zc = max(p for p in range(2, m) if all(p % d for d in range(2, int(p ** 0.5) + 1)))
    qbar = n_zc * (u + 1) / 31.0
    q = np.floor(qbar + 0.5) + v * (-1.0) ** np.floor(2 * qbar)
    mm = np.arange(n_zc, dtype=np.float64)
    xq = np.exp(-1j * np.pi * q * mm * (mm + 1) / n_zc)
    return xq[np.arange(m) % n_zc]


def test_unsupported_lengths_are_refused_without_a_gpu():
    for m in (6, 12, 18, 24, 5000):
        rc = _lib.lib.srsce_gen_pilots(None, 0, m, 1, 1, None, None, None)
        assert rc == _lib.ERR_UNSUPPORTED, (m, rc, _lib.last_error())
    assert _lib.lib.srsce_gen_pilots(None, 0, 36, 0, 1, None, None, None) == _lib.ERR_INVALID


@pytest.mark.gpu
@pytest.mark.parametrize("m", [30, 36, 48, 72, 144, 816, 819, 1638])
def test_sequences_match_the_specification_formulas(m):
    from srsran_ce_pytorch_b200 import sequences
    u = torch.tensor([0, 1, 7, 13, 29, 29], dtype=torch.int32, device="cuda")
    v = torch.tensor([0, 0, 1 if m >= 72 else 0, 0, 0, 1 if m >= 72 else 0], dtype=torch.int32, device="cuda")
    p = sequences.low_papr_pilots(u, m, 3, 4, v).cpu().numpy()
    assert p.shape == (6, m, 3, 4)
    occ = (-1.0) ** np.arange(m)
    for b in range(6):
        want = _spec_sequence(int(u[b]), int(v[b]), m)
        # float32 storage of a unit-modulus float64 value: 6e-8; the spec formula evaluated in float64 carries ~1e-10 of
        # phase error for the largest arguments (pi q m (m+1) / N_zc ~ 1e7), the kernel's integer-reduced phase none
        assert np.abs(p[b, :, 0, 0] - want).max() < 2e-7, (m, b)
        assert np.abs(p[b, :, 2, 2] - want).max() < 2e-7
        assert np.abs(p[b, :, 1, 1] - occ * want).max() < 2e-7 and np.array_equal(p[b, :, 1, 1], p[b, :, 0, 3])
        assert np.abs(np.abs(p[b]) - 1.0).max() < 2e-7                      # constant modulus
    if m >= 36:                                                             # Zadoff-Chu core: ideal periodic autocorrelation
        n_zc = max(q for q in range(2, m) if all(q % d for d in range(2, int(q ** 0.5) + 1)))
        x = p[1, :n_zc, 0, 0].astype(np.complex128)
        ac = np.fft.ifft(np.abs(np.fft.fft(x)) ** 2)
        assert np.abs(ac[1:]).max() < 2e-5 * abs(ac[0])


@pytest.mark.gpu
def test_generated_pilots_through_the_estimator():
    """Pilots generated on the device, a flat per-layer channel, no noise: the estimator returns the channel (comb-4,
    4 ports = 2 combs x the OCC pair made of cyclic shifts 0 and pi), RSRP = mean |h|^2, noise ~ 0."""
    from srsran_ce_pytorch_b200 import estimator, sequences, synth
    n_prb, L = 48, 4
    hop1 = synth.make_hop(n_prb, 0, n_prb, [13], "comb4", 2)
    cfg = synth.make_config(30e3, "filter", True)
    n_re = n_prb * 3
    pil = sequences.low_papr_pilots(torch.tensor([5, 11], dtype=torch.int32, device="cuda"), n_re, 1, L)
    h = torch.tensor([[0.8 + 0.3j, -0.2 + 0.9j, 0.5 - 0.5j, -0.7 - 0.1j], [0.1 + 1.0j, 0.6 + 0.2j, -0.4 + 0.4j, 0.9 - 0.3j]],
                     dtype=torch.complex64, device="cuda")
    rx = torch.zeros((2, 12 * n_prb, 14), dtype=torch.complex64, device="cuda")
    pos = [torch.arange(0, 12 * n_prb, 4, device="cuda"), torch.arange(2, 12 * n_prb, 4, device="cuda")]   # comb offsets 0 and 2
    for b in range(2):
        for g in range(2):
            rx[b, pos[g], 13] = pil[b, :, 0, 2 * g] * h[b, 2 * g] + pil[b, :, 0, 2 * g + 1] * h[b, 2 * g + 1]
    out, sc = estimator.srs_channel_estimator_batched(rx, pil, 1.0, hop1, synth.no_hop(), cfg, pilot_share=1)
    for b in range(2):
        err = (out[b] - h[b][None, None, :]).abs().max()
        assert float(err) < 2e-5, (b, float(err))
        assert abs(float(sc[b, 1]) - float((h[b].abs() ** 2).mean())) < 1e-5          # rsrp
        assert float(sc[b, 0]) < 1e-9 and float(sc[b, 3]) == 0.0                     # noise, time alignment
