"""Ensemble statistics of epicurves: the recipe the reference's own tooling uses to compare runs
(engine/plot/models/EpiCurves.py:25-40: truncate every run to the shortest, then mean and standard deviation per hour and
compartment; engine/plot/collate_all_simulations.py) plus the peak-infection summary BASELINE.json's north_star names.

Pure numpy on Counts rows [hour, S, E, I, H, R, D]; used by tests/test_ensemble_gpu.py and by users collating runs.
"""
import numpy as np

COMPARTMENTS = ("susceptible", "exposed", "infected", "hospitalized", "recovered", "deceased")


def pad_to_hours(rows, hours):
    """A run that stopped early (Epidemiology::stop_simulation: no exposed / infected / hospitalized agents left) stays in its
    final state: repeat the last row so every run of an ensemble has `hours` rows."""
    rows = np.asarray(rows)
    if len(rows) >= hours:
        return rows[:hours]
    tail = np.repeat(rows[-1:], hours - len(rows), axis=0).copy()
    tail[:, 0] = np.arange(int(rows[-1, 0]) + 1, int(rows[-1, 0]) + 1 + len(tail))
    return np.concatenate([rows, tail], axis=0)


def make_number_of_rows_equal(runs):
    """EpiCurves.py:25-27: truncate every run to the shortest one."""
    n = min(len(r) for r in runs)
    return np.stack([np.asarray(r)[:n] for r in runs]).astype(np.float64)


def mean_and_std(runs):
    """EpiCurves.py:30-40.  Returns (hours[n], mean[n, 6], std[n, 6]) over the runs (population std, numpy default like pandas' .std(ddof=0)
    is NOT what pandas uses: pandas DataFrame.std is ddof=1, and so is this)."""
    a = make_number_of_rows_equal(runs)
    return a[0, :, 0], a[:, :, 1:].mean(axis=0), a[:, :, 1:].std(axis=0, ddof=1)


def peak_infected(rows):
    """(peak magnitude, hour of the peak) of the infected column (I, not counting hospitalized)."""
    rows = np.asarray(rows)
    k = int(np.argmax(rows[:, 3]))
    return float(rows[k, 3]), float(rows[k, 0])


def compare(candidate_runs, reference_runs, z=1.96):
    """Candidate ensemble mean against the reference ensemble's band (mean +- z * std per hour and compartment), plus the
    standardised difference of the two means.  Returns a dict of summary numbers."""
    _, mc, sc = mean_and_std(candidate_runs)
    _, mr, sr = mean_and_std(reference_runs)
    n = min(len(mc), len(mr))
    mc, sc, mr, sr = mc[:n], sc[:n], mr[:n], sr[:n]
    nc, nr = len(candidate_runs), len(reference_runs)
    inside = np.abs(mc - mr) <= z * sr + 0.5  # + 0.5: counts are integers; a zero-variance hour must still admit rounding
    se = np.sqrt(sc ** 2 / nc + sr ** 2 / nr)
    zscore = np.where(se > 0, np.abs(mc - mr) / np.maximum(se, 1e-12), 0.0)
    pk_c = np.array([peak_infected(r) for r in candidate_runs])
    pk_r = np.array([peak_infected(r) for r in reference_runs])

    def welch(a, b):
        s = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
        return float(abs(a.mean() - b.mean()) / s) if s > 0 else 0.0

    return {
        "hours": n,
        "fraction_inside_band": float(inside.mean()),
        "inside_band_by_compartment": {c: float(inside[:, k].mean()) for k, c in enumerate(COMPARTMENTS)},
        "max_z_of_means": float(zscore.max()),
        "fraction_z_below_3": float((zscore < 3.0).mean()),
        "peak_magnitude": {"candidate": float(pk_c[:, 0].mean()), "reference": float(pk_r[:, 0].mean()), "reference_std": float(pk_r[:, 0].std(ddof=1)),
                           "z": welch(pk_c[:, 0], pk_r[:, 0])},
        "peak_hour": {"candidate": float(pk_c[:, 1].mean()), "reference": float(pk_r[:, 1].mean()), "reference_std": float(pk_r[:, 1].std(ddof=1)),
                      "z": welch(pk_c[:, 1], pk_r[:, 1])},
    }
