"""Population::Csv: Grid::read_population + Citizen::from_record
(engine/src/geography/grid.rs:194-231, citizen/mod.rs:155-180, citizen/population_record.rs:23-43).
CPU: the product's reader + factory (host code behind epi_population_size / epi_build_population) against the oracle's
independent reader + restatement, and against Python's csv module.  GPU: a run from a CSV population, bit-exact vs the oracle."""
import csv
import os

import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200.engine import STATE_FIELDS, Engine, EpiError, build_population, make_config, population_size

AGES = ["0-4", "25-29", "60-64", "80+"]


def write_population(path, n, seed=0, extra_cols=True, crlf=False):
    rng = np.random.default_rng(seed)
    working = rng.random(n) < 0.6
    pt = rng.random(n) < 0.3  # not gated on `working`: Citizen::from_record takes the column as it is
    with open(path, "w", newline="") as f:
        w = csv.writer(f, lineterminator="\r\n" if crlf else "\n")
        w.writerow(["ind", "age", "sex", "working", "pub_transport"] if extra_cols else ["working", "ind", "pub_transport", "age"])
        for i in range(n):
            age = AGES[i % 4] if i % 7 else 'sixty, "or so"'  # a quoted field with a comma and quotes
            if extra_cols:
                w.writerow([i, age, "MF"[i % 2], str(bool(working[i])), str(bool(pt[i]))])
            else:
                w.writerow([str(bool(working[i])), i, str(bool(pt[i])), age])
    return working, pt


@pytest.mark.parametrize("extra_cols,crlf", [(True, False), (False, True)])
def test_csv_population_factory_matches_oracle(tmp_path, extra_cols, crlf):
    path = tmp_path / "pop.csv"
    working, pt = write_population(path, 4321, seed=5, extra_cols=extra_cols, crlf=crlf)
    kw = dict(grid_size=150, exposed=4, mild=2, population_csv=path, n_agents=7)  # n_agents is ignored
    cfg = make_config(**kw)
    assert population_size(cfg) == 4321
    ours = build_population(cfg, seed=9)
    orc = O.OracleEngine(O.make_config(**kw), seed=9).get_state()
    for f in STATE_FIELDS:
        assert (ours[f] == orc[f]).all(), f"{f} differs"
    # record c describes the citizen of house c % H (grid.rs:205-208): same (house, working, pub_transport) triples
    H = 30 * 75  # houses of G = 150: floor(60 / 2) x floor(151 / 2)
    ws = (ours["st"] >> 13) & 3
    got = sorted(zip(ours["home"].tolist(), (ws != 3).tolist(), (((ours["st"] >> 9) & 1) == 1).tolist()))
    want = sorted(zip((np.arange(4321) % H).tolist(), working.tolist(), pt.tolist()))
    assert got == want
    assert ((ours["st"] & 7) == 1).sum() == 4  # starting infections still apply (grid.rs:226)


def test_csv_errors(tmp_path):
    def size_of(text):
        p = tmp_path / "bad.csv"
        p.write_text(text)
        return population_size(make_config(grid_size=100, population_csv=p))

    assert size_of("ind,age,working,pub_transport\n1,20-24,True,False\n\n2,80+,False,False\n") == 2
    with pytest.raises(EpiError, match="True or False"):  # population_record.rs:34-43
        size_of("ind,age,working,pub_transport\n1,20-24,true,False\n")
    with pytest.raises(EpiError, match="missing field `pub_transport`"):
        size_of("ind,age,working\n1,20-24,True\n")
    with pytest.raises(EpiError, match="fields"):
        size_of("ind,age,working,pub_transport\n1,20-24,True\n")
    with pytest.raises(EpiError, match="ind"):
        size_of("ind,age,working,pub_transport\nx,20-24,True,False\n")
    with pytest.raises(EpiError, match="Could not read population file"):  # grid.rs:202
        population_size(make_config(grid_size=100, population_csv=tmp_path / "absent.csv"))
    p = tmp_path / "big.csv"
    write_population(p, 2000)
    with pytest.raises(EpiError, match="Cannot accommodate citizens into homes"):  # grid.rs:216-223: G = 60 has 12 x 30 houses
        build_population(make_config(grid_size=60, population_csv=p))
    p.write_text("ind,age,working,pub_transport\n")
    with pytest.raises(EpiError, match="no records"):
        build_population(make_config(grid_size=100, population_csv=p))


@pytest.mark.gpu
def test_csv_population_run_matches_oracle(tmp_path):
    path = tmp_path / "pop.csv"
    write_population(path, 30000, seed=2)
    kw = dict(grid_size=450, exposed=300, asym=30, mild=30, severe=30, population_csv=path, lockdown=(500, 0.1))
    with Engine(make_config(**kw), seed=4) as gpu:
        orc = O.OracleEngine(O.make_config(**kw), seed=4)
        assert gpu.population == orc.population == 30000
        for hour in range(1, 60):
            cg, co = gpu.step(hour), orc.step(hour)
            assert (cg == co).all(), f"hour {hour}: gpu {cg} != oracle {co}"
        a, b = gpu.get_state(), orc.get_state()
        for f in STATE_FIELDS:
            assert (a[f] == b[f]).all(), f"state field {f} differs"


@pytest.mark.gpu
def test_engine_app_cli_with_a_csv_population(tmp_path):
    """engine-app -c <config with population.Csv>: the file path is relative to the working directory, as in the reference
    (engine/config/pune.json: "config/pune_population.csv"); the epicurve CSV equals the oracle's rows for the same seed."""
    import glob
    import json
    import subprocess

    from epirust_b200 import build as B

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    write_population(tmp_path / "pop.csv", 8000, seed=11)
    cfg = json.load(open(os.path.join(root, "tests", "golden", "default_config.json")))
    cfg["population"] = {"Csv": {"file": "pop.csv", "cols": ["age", "sex", "working", "pub_transport"]}}
    cfg["hours"] = 120
    cfg["starting_infections"] = {"infected_mild_asymptomatic": 10, "infected_mild_symptomatic": 10, "infected_severe": 10, "exposed": 100}
    (tmp_path / "csv_pop.json").write_text(json.dumps(cfg))
    B.build()
    r = subprocess.run([B.APP, "-c", "csv_pop.json", "-o", str(tmp_path), "--seed", "6"], capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    (out,) = glob.glob(str(tmp_path / "output" / "simulation_0_*[0-9].csv"))
    got = np.array([row for row in csv.reader(open(out))][1:], dtype=np.uint32)
    kw = dict(grid_size=250, hours=120, exposed=100, asym=10, mild=10, severe=10, lockdown=(100, 0.1), population_csv=tmp_path / "pop.csv")
    rows_o, _, _ = O.oracle_run(O.make_config(**kw), seed=6, mode="keyed", threads=2)
    assert got.shape == rows_o.shape and (got == rows_o).all()
    assert (got[:, 1:].sum(axis=1) == 8000).all()
