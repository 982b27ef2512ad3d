"""Generates tests/golden/ensemble_stream.npz: the reference-like ensemble (oracle STREAM mode: sequential per-thread RNG streams consumed
like the reference consumes thread_rng, hash-order phase B) for the ensemble workload of tests/test_ensemble_gpu.py.

    python tests/golden/make_ensemble_fixture.py        (about two minutes on 8 cores)

The Rust reference cannot be run in this image, so this is the oracle's ensemble, not EpiRust's."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_ffi as O  # noqa: E402
from epirust_b200 import ensemble as E  # noqa: E402

WORKLOAD = dict(n_agents=10000, grid_size=250, hours=1080, exposed=50, lockdown=(100, 0.1))  # default.json with 50 initial exposed
SEEDS = list(range(101, 165))  # 64 runs


def stream_run(seed):
    return E.pad_to_hours(O.oracle_run(O.make_config(**WORKLOAD), seed=seed, mode="stream", threads=1)[0], WORKLOAD["hours"] - 1)


if __name__ == "__main__":
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        runs = list(ex.map(stream_run, SEEDS))
    hours, mean, std = E.mean_and_std(runs)
    peaks = np.array([E.peak_infected(r) for r in runs])
    np.savez_compressed(os.path.join(HERE, "ensemble_stream.npz"), hours=hours.astype(np.uint32), mean=mean.astype(np.float32), std=std.astype(np.float32),
                        peaks=peaks.astype(np.float32), seeds=np.array(SEEDS), workload=np.array(repr(WORKLOAD)))
    print("runs", len(runs), "peak I mean", peaks[:, 0].mean(), "+-", peaks[:, 0].std(ddof=1), "peak hour", peaks[:, 1].mean(), "+-", peaks[:, 1].std(ddof=1))
