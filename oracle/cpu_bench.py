"""CPU timing of the oracle port over all host cores (the ``cpu_baseline`` leg and the
``--impl reference`` arm of bench.py).  TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never on the
product path.  Imports numpy and the oracle only, so that spawned workers start quickly.

The reference itself (/root/reference/src/ce_rule_tensorized.py) is Python on torch eager ops and
does not travel to the GPU box; the numpy port computes the same function and was measured 2.9x
FASTER than the reference at BASELINE config 2 in the build container (3.8 ms vs 11.0 ms per
unit, one thread), so it is a generous stand-in for the reference's CPU path.
"""
from __future__ import annotations

import os
import time

import numpy as np

_STATE = {}


def _init(rx, pilots, share, beta, hop1, hop2, cfg, fill="tensorized"):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import ce_oracle
    _STATE.update(rx=rx, pilots=pilots, share=share, beta=beta, hop1=hop1, hop2=hop2, cfg=cfg, fill=fill,
                  est=ce_oracle.estimate)
    ce_oracle.estimate(rx[0], pilots[0], beta, hop1, hop2, cfg, fill=fill)      # warm caches / imports


def _work(args):
    first, count = args
    s = _STATE
    n = s["rx"].shape[0]
    acc = 0.0
    for k in range(first, first + count):
        b = k % n
        out = s["est"](s["rx"][b], s["pilots"][b // s["share"]], s["beta"], s["hop1"], s["hop2"], s["cfg"], fill=s["fill"])
        acc += float(out[2])
    return acc


class OraclePool:
    """``cores`` single-threaded workers, each holding the same small set of distinct units."""

    def __init__(self, rx: np.ndarray, pilots: np.ndarray, share: int, beta: float, hop1, hop2, cfg,
                 cores: int | None = None, fill: str = "tensorized"):
        import multiprocessing as mp
        self.cores = cores or len(os.sched_getaffinity(0))
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_init, initargs=(rx, pilots, share, beta, hop1, hop2, cfg, fill))
        self.pool.map(_work, [(0, 1)] * self.cores)                  # every worker is up and warm

    def run(self, units_per_worker: int) -> float:
        """Process cores * units_per_worker units; returns wall seconds."""
        t0 = time.perf_counter()
        self.pool.map(_work, [(w * units_per_worker, units_per_worker) for w in range(self.cores)], chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
