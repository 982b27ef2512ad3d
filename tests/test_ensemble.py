"""Ensemble statistics (the reference's own comparison recipe, engine/plot/models/EpiCurves.py:25-40) and the statistical
equivalence of the keyed / lowest-id convention with the reference-like STREAM convention, on the CPU oracle."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

import oracle_ffi as O
from epirust_b200 import ensemble as E

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WORKLOAD = dict(n_agents=10000, grid_size=250, hours=1080, exposed=50, lockdown=(100, 0.1))


def load_fixture():
    z = np.load(os.path.join(GOLDEN, "ensemble_stream.npz"))
    assert str(z["workload"]) == repr(WORKLOAD)
    return z


def assert_inside_reference_band(runs, z=1.96):
    """North-star bar: per-hour compartment means, peak-infection magnitude and peak hour of the candidate ensemble fall inside the
    reference ensemble's 95 % interval (mean +- 1.96 std of the committed STREAM ensemble)."""
    fx = load_fixture()
    _, mean, _ = E.mean_and_std(runs)
    n = min(len(mean), len(fx["mean"]))
    out = np.abs(mean[:n] - fx["mean"][:n]) > z * fx["std"][:n] + 0.5
    assert not out.any(), f"{out.sum()} (hour, compartment) means outside the band, first at hour {fx['hours'][np.nonzero(out.any(axis=1))[0][0]]}"
    pk = np.array([E.peak_infected(r) for r in runs]).mean(axis=0)
    ref_mean, ref_std = fx["peaks"].mean(axis=0), fx["peaks"].std(axis=0, ddof=1)
    assert abs(pk[0] - ref_mean[0]) <= z * ref_std[0], f"peak infected {pk[0]} vs {ref_mean[0]} +- {ref_std[0]}"
    assert abs(pk[1] - ref_mean[1]) <= z * ref_std[1], f"peak hour {pk[1]} vs {ref_mean[1]} +- {ref_std[1]}"


def test_mean_std_recipe_truncates_to_the_shortest_run():
    a = np.array([[1, 10, 0, 0, 0, 0, 0], [2, 9, 1, 0, 0, 0, 0], [3, 8, 1, 1, 0, 0, 0]])
    b = np.array([[1, 10, 0, 0, 0, 0, 0], [2, 7, 3, 0, 0, 0, 0]])
    hours, mean, std = E.mean_and_std([a, b])
    assert hours.tolist() == [1, 2] and mean[1].tolist() == [8, 2, 0, 0, 0, 0]
    assert np.allclose(std[1], [np.std([9, 7], ddof=1), np.std([1, 3], ddof=1), 0, 0, 0, 0])
    padded = E.pad_to_hours(b, 4)
    assert padded[:, 0].tolist() == [1, 2, 3, 4] and padded[3, 1:].tolist() == [7, 3, 0, 0, 0, 0]
    assert E.peak_infected(a) == (1.0, 3.0)


def test_post_processing_recipes_of_the_reference_plot_scripts(tmp_path):
    # engine/plot/collate_all_simulations.py (EpiCurves.to_csv), update_total_infections.py, merge_regions_data.py
    a = np.array([[1, 10, 0, 0, 0, 0, 0], [2, 9, 1, 0, 0, 0, 0], [3, 8, 1, 1, 0, 0, 0]])
    b = np.array([[1, 10, 0, 0, 0, 0, 0], [2, 7, 3, 0, 0, 0, 0]])
    out = tmp_path / "collated_simulation.csv"
    E.collate_to_csv([a, b], out)
    lines = out.read_text().splitlines()
    assert lines[0] == "susceptible,susceptible_std,exposed,exposed_std,infected,infected_std,hospitalized,hospitalized_std,recovered,recovered_std,deceased,deceased_std,hour"
    row2 = [float(v) for v in lines[2].split(",")]
    assert row2[:4] == [8.0, 1.0, 2.0, 1.0] and row2[-1] == 2 and len(lines) == 3  # population std (numpy .std()), shortest run
    rows = np.array([[h, 0, 0, i, hsp, r, d] for h, (i, hsp, r, d) in enumerate([(1, 0, 0, 0), (3, 1, 0, 0), (2, 1, 2, 1), (0, 0, 5, 2)], 1)])
    t = E.with_total_infected(rows, ma_window=2)
    assert t["totalinfected"].tolist() == [1, 4, 6, 7]
    assert np.isnan(t["ma_infected"][0]) and t["ma_infected"][1:].tolist() == [2.0, 2.5, 1.0] and t["ma_deceased"][3] == 1.5
    m = E.merge_regions([a, b])
    assert m[:, 1].tolist() == [20, 16, 8] and m[:, 0].tolist() == [2, 4, 3]  # the script sums the hour column too
    p = tmp_path / "simulation_0_x.csv"
    p.write_text("hour,susceptible,exposed,infected,hospitalized,recovered,deceased\n1,10,0,0,0,0,0\n2,9,1,0,0,0,0\n")
    assert E.read_rows(p).tolist() == [[1, 10, 0, 0, 0, 0, 0], [2, 9, 1, 0, 0, 0, 0]]


def test_compare_flags_a_shifted_ensemble():
    rng = np.random.default_rng(0)
    base = np.zeros((20, 50, 7))
    base[:, :, 0] = np.arange(1, 51)
    base[:, :, 3] = 100 + rng.normal(0, 5, (20, 50))
    same = base.copy()
    same[:, :, 3] = 100 + rng.normal(0, 5, (20, 50))
    shifted = base.copy()
    shifted[:, :, 3] += 40
    assert E.compare(list(same), list(base))["fraction_inside_band"] == 1.0
    assert E.compare(list(shifted), list(base))["inside_band_by_compartment"]["infected"] == 0.0


def test_keyed_lowest_id_ensemble_is_inside_the_stream_band():
    """The draw-slot / lowest-id convention the GPU uses (oracle KEYED mode) against the committed reference-like ensemble."""
    def keyed(seed):
        return E.pad_to_hours(O.oracle_run(O.make_config(**WORKLOAD), seed=seed, mode="keyed", threads=1)[0], WORKLOAD["hours"] - 1)

    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        runs = list(ex.map(keyed, range(1, 33)))
    assert_inside_reference_band(runs)


def test_keyed_and_stream_ensembles_agree_with_all_three_interventions():
    """A second workload for the same claim, with every intervention on the path active (lockdown, hospital build-up, vaccination)
    and a denser start: 32 KEYED runs (the GPU's convention) against 32 fresh STREAM runs (the reference-like convention), compared
    with the recipe of engine/plot/models/EpiCurves.py -- per-hour means inside the STREAM ensemble's 95 % band, and the peaks
    of the two ensembles statistically indistinguishable (Welch z < 3)."""
    wl = dict(n_agents=20000, grid_size=350, hours=720, exposed=200, lockdown=(400, 0.1), hospital=100, vaccinate=((240, 0.2),))

    def run(args):
        mode, seed = args
        return E.pad_to_hours(O.oracle_run(O.make_config(**wl), seed=seed, mode=mode, threads=1)[0], wl["hours"] - 1)

    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        runs = list(ex.map(run, [("keyed", s) for s in range(1, 33)] + [("stream", s) for s in range(201, 233)]))
    keyed, stream = runs[:32], runs[32:]
    c = E.compare(keyed, stream)
    assert c["fraction_inside_band"] > 0.995, c
    assert c["peak_magnitude"]["z"] < 3.0 and c["peak_hour"]["z"] < 3.0, c
    assert c["peak_magnitude"]["reference"] > 1000  # a real epidemic, not a fizzle
