"""Whole-run statistical parity (BASELINE.json north_star): a 32-seed ensemble of GPU runs (Philox keyed draws, lowest-id conflict
priority) against the reference-like ensemble (oracle STREAM mode: sequential RNG streams consumed like thread_rng, hash-order
phase B) -- per-hour compartment means, peak-infection magnitude and peak hour inside the reference ensemble's 95 % interval."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200 import ensemble as E
from epirust_b200.engine import make_config, run_standalone
from test_ensemble import WORKLOAD, assert_inside_reference_band

pytestmark = pytest.mark.gpu


def gpu_ensemble(seeds):
    return [E.pad_to_hours(run_standalone(make_config(**WORKLOAD), seed=s)[0], WORKLOAD["hours"] - 1) for s in seeds]


def test_gpu_ensemble_inside_the_committed_reference_band():
    assert_inside_reference_band(gpu_ensemble(range(1, 33)))


def test_gpu_ensemble_against_a_fresh_stream_ensemble():
    """Same comparison against a STREAM ensemble computed on this box with other seeds, plus the standardised difference of means."""
    def stream(seed):
        return E.pad_to_hours(O.oracle_run(O.make_config(**WORKLOAD), seed=seed, mode="stream", threads=1)[0], WORKLOAD["hours"] - 1)

    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        ref = list(ex.map(stream, range(1001, 1033)))
    rep = E.compare(gpu_ensemble(range(41, 73)), ref)
    assert rep["fraction_inside_band"] == 1.0, rep
    assert rep["fraction_z_below_3"] >= 0.99, rep
    assert rep["peak_magnitude"]["z"] < 3.5 and rep["peak_hour"]["z"] < 3.5, rep
    pm, ph = rep["peak_magnitude"], rep["peak_hour"]
    assert abs(pm["candidate"] - pm["reference"]) <= 1.96 * pm["reference_std"], rep
    assert abs(ph["candidate"] - ph["reference"]) <= 1.96 * ph["reference_std"], rep
