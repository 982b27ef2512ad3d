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


def mean_and_std(runs, ddof=1):
    """EpiCurves.py:30-40.  Returns (hours[n], mean[n, 6], std[n, 6]) over the runs.  The reference's collation takes numpy's
    population standard deviation (`collated_columns.std(axis=0)` on an ndarray: ddof=0) -- pass ddof=0 for that; the
    default ddof=1 is the sample estimate the confidence bands of `compare` need."""
    a = make_number_of_rows_equal(runs)
    return a[0, :, 0], a[:, :, 1:].mean(axis=0), a[:, :, 1:].std(axis=0, ddof=ddof)


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


# ---- the reference's post-processing scripts on Counts rows (engine/plot/*.py) ------------------------------------------------
def read_rows(csv_path):
    """An epicurve CSV (`hour,susceptible,exposed,infected,hospitalized,recovered,deceased`, listeners/csv_service.rs:44-71) as rows[n, 7]."""
    with open(csv_path) as f:
        header = f.readline().strip().split(",")
        if header != ["hour"] + list(COMPARTMENTS):
            raise ValueError(f"{csv_path}: not an epicurve CSV (header {header})")
        return np.loadtxt(f, delimiter=",", dtype=np.int64, ndmin=2)


def collate_to_csv(runs, output_path):
    """collate_all_simulations.py --output-path (EpiCurves.to_csv, EpiCurves.py:87-97): per compartment the mean and the
    population standard deviation over the runs, columns `<name>,<name>_std,...,hour` with hour = row index + 1."""
    _, mean, std = mean_and_std(runs, ddof=0)
    with open(output_path, "w") as f:
        f.write(",".join(f"{c},{c}_std" for c in COMPARTMENTS) + ",hour\n")
        for k in range(len(mean)):
            f.write(",".join(f"{float(mean[k, j])!r},{float(std[k, j])!r}" for j in range(len(COMPARTMENTS))) + f",{k + 1}\n")


def with_total_infected(rows, ma_window=0):
    """update_total_infections.py:30-36: totalinfected = infected + recovered + deceased + hospitalized, and -- for a window > 0 --
    the moving averages of infected and deceased (NaN until the window is full, like pandas' rolling().mean())."""
    rows = np.asarray(rows, dtype=np.float64)
    out = {"totalinfected": rows[:, 3] + rows[:, 5] + rows[:, 6] + rows[:, 4]}
    if ma_window > 0:
        for name, col in (("ma_infected", 3), ("ma_deceased", 6)):
            c = np.concatenate([[0.0], np.cumsum(rows[:, col])])
            ma = np.full(len(rows), np.nan)
            if len(rows) >= ma_window:
                ma[ma_window - 1:] = (c[ma_window:] - c[:-ma_window]) / ma_window
            out[name] = ma
    return out


def merge_regions(region_rows):
    """merge_regions_data.py:33-34: element-wise sum of the regions' tables, a shorter table counting as zeros
    (`DataFrame.add(fill_value=0.0)`; the hour column is summed as well, exactly like the script does)."""
    n = max(len(r) for r in region_rows)
    total = np.zeros((n, np.asarray(region_rows[0]).shape[1]), np.float64)
    for r in region_rows:
        r = np.asarray(r, dtype=np.float64)
        total[: len(r)] += r
    return total
