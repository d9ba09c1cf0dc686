"""The numpy oracle (oracle/ce_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, generated in the build container by oracle/gen_golden.py)."""
import numpy as np
import pytest

from oracle import ce_oracle
from tests import tolerances as tol
from tests.golden_io import case_names, load_case

NAMES = case_names()


def test_golden_set_is_present():
    assert len(NAMES) >= 40
    for must in ("case0_shape", "case4_shape", "case8_shape", "cfg2_273prb_comb4_L4_S1"):
        assert must in NAMES


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    c = load_case(name)
    n_checked = 0
    for variant in ("tensorized", "baseline", "cnn"):
        if f"{variant}_grid" not in c["ref"]:
            continue
        if variant == "cnn" and c["rx"].shape[0] > 1000:
            continue                       # 409 inpaint iterations x 273 PRB: covered once below
        diag = {}
        out = ce_oracle.estimate(c["rx"], c["pilots"], c["beta"], c["hop1"], c["hop2"], c["cfg"],
                                 fill=variant, diagnostics=diag)
        mx, nmse = tol.grid_errors(out[0], c["ref"][f"{variant}_grid"])
        assert out[0].dtype == np.complex64
        assert mx <= tol.GRID_MAX_REL_TIGHT and nmse <= tol.GRID_NMSE, (name, variant, mx, nmse)
        tol.check_scalars(out[1:], c["ref"][f"{variant}_scalars"], float(c["cfg"].scs),
                          diag["ta_margin"], tag=f"{name}/{variant}")
        n_checked += 1
    assert n_checked >= 1


def test_oracle_cnn_full_band():
    c = load_case("cfg2_273prb_comb4_L4_S1")
    out = ce_oracle.estimate(c["rx"], c["pilots"], c["beta"], c["hop1"], c["hop2"], c["cfg"], fill="cnn")
    mx, nmse = tol.grid_errors(out[0], c["ref"]["cnn_grid"])
    assert mx <= tol.GRID_MAX_REL_TIGHT and nmse <= tol.GRID_NMSE


def test_reference_variants_agree_in_golden():
    """baseline == tensorized to fp32 rounding -- guards the fixtures themselves."""
    for name in NAMES:
        c = load_case(name)
        if "baseline_grid" in c["ref"]:
            mx, _ = tol.grid_errors(c["ref"]["baseline_grid"], c["ref"]["tensorized_grid"])
            assert mx <= 1e-6, (name, mx)


def test_rc_taps_table():
    """Tap counts / virtual-pilot counts of SURVEY Appendix A.3."""
    expect = {  # REs per PRB -> [(taps, n_virtual) for nPRB = 1, 2, >=3]
        12: [(11, 12), (21, 10), (31, 12)], 6: [(5, 6), (11, 5), (15, 7)], 4: [(3, 4), (7, 3), (11, 5)],
        3: [(3, 3), (5, 2), (7, 3)], 2: [(1, 2), (3, 1), (5, 2)], 1: [(1, 1), (1, 0), (3, 1)],
    }
    for r, rows in expect.items():
        for n_prb, (n_taps, nv) in zip((1, 2, 5), rows):
            t = ce_oracle.rc_taps(12 // r, min(3, n_prb))
            assert t.size == n_taps and abs(t.sum() - 1.0) < 1e-12
            assert ce_oracle.n_virtual(t.size, n_prb, r) == nv
    t = ce_oracle.rc_taps(4, 3)
    assert np.allclose(t, [-0.0554, 0.08565, 0.2822, 0.37511, 0.2822, 0.08565, -0.0554], atol=6e-5)


def test_errors_match_reference_behaviour():
    c = load_case("comb2_L2")
    c["cfg"].Smoothing = "bogus"
    with pytest.raises(ValueError):
        ce_oracle.estimate(c["rx"], c["pilots"], c["beta"], c["hop1"], c["hop2"], c["cfg"])
    c = load_case("2hop_L2_comb2")
    c["hop2"].DMRSsymbols = c["hop1"].DMRSsymbols.clone()
    with pytest.raises(AssertionError):
        ce_oracle.estimate(c["rx"], c["pilots"], c["beta"], c["hop1"], c["hop2"], c["cfg"])
