"""Error structure of one case of the random campaign (tests/test_gpu_fuzz.py): per-symbol / per-layer NMSE and the scalars.
    FUZZ_N=400 FUZZ_SEED=4101 python tools/fuzz_case_probe.py 293 [unit]"""
import importlib
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from oracle import ce_oracle
from srsran_ce_pytorch_b200 import estimator, synth

fz = importlib.import_module("tests.test_gpu_fuzz")
k = int(sys.argv[1])
c = fz.CASES[k]
print(c)
n_cdm = (c["layers"] + 1) // 2
a0, an = c["alloc"]
hop1 = synth.make_hop(c["grid"], c["start"], c["n_prb"], c["s1"], c["cols"], n_cdm, start_symbol=a0, n_alloc=an)
hop2 = synth.make_hop(c["grid"], c["start2"], c["n_prb"], c["s2"], c["cols"], n_cdm, start_symbol=a0, n_alloc=an) if c["s2"] else synth.no_hop()
cfg = synth.make_config(30e3, c["sm"], c["cfo"])
data = synth.make_units(hop1, hop2, cfg, c["layers"], 3, pilot_share=1, seed=500 + k, device="cuda", beta=c["beta"], snr_db=c["snr"])
out, sc = estimator.srs_channel_estimator_batched(data["rx"], data["pilots"], c["beta"], hop1, hop2, cfg, pilot_share=1)
out, sc = out.cpu().numpy().astype(np.complex128), sc.cpu().numpy()
rx, pil = data["rx"].cpu().numpy(), data["pilots"].cpu().numpy()
for b in range(3) if len(sys.argv) < 3 else [int(sys.argv[2])]:
    want = ce_oracle.estimate(rx[b], pil[b], c["beta"], hop1, hop2, cfg)
    w = want[0].astype(np.complex128)
    d = out[b] - w
    print(f"unit {b}: NMSE {np.sum(abs(d)**2) / np.sum(abs(w)**2):.3e}  scalars got {sc[b]}  want {want[1:]}")
    for t in range(w.shape[1]):
        den = np.sum(abs(w[:, t]) ** 2)
        if den > 0:
            ratio = np.sum(out[b][:, t] * np.conj(w[:, t])) / den
            print(f"   symbol {t:2d}: NMSE {np.sum(abs(d[:, t])**2) / den:.3e}   got/want = 1 + ({ratio.real - 1:+.3e}) + j({ratio.imag:+.3e})")
    for l in range(w.shape[2]):
        den = np.sum(abs(w[:, :, l]) ** 2)
        print(f"   layer {l}: NMSE {np.sum(abs(d[:, :, l])**2) / den:.3e}")
if len(sys.argv) > 3:                                    # dump the unit for offline analysis
    b = int(sys.argv[2])
    np.savez(sys.argv[3], rx=rx[b], pil=pil[b], out=out[b].astype(np.complex64), sc=sc[b])
