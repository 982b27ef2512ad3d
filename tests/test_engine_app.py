"""The outer boundary: the engine-app command line (engine-app/src/main.rs:60-177), the multi-region Configuration
(common/src/config/configuration.rs:28-117) and the listeners' output files.  CPU-only checks here (host-only entries of the
C ABI); the runs themselves are in the -m gpu tests."""
import json
import os
import subprocess

import numpy as np
import pytest

from epirust_b200 import build as B
from epirust_b200.engine import Configuration, EpiError, write_outputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def app():
    B.build()
    assert os.path.exists(B.APP)
    return B.APP


def run(app, *args, env=None):
    return subprocess.run([app, *args], capture_output=True, text=True, timeout=120, env=env)


def test_help_lists_the_reference_flags(app):
    r = run(app, "--help")
    assert r.returncode == 0
    for flag in ("-c, --config <FILE>", "-m, --mode <MODE>", "-i, --id <ID>", "-t, --threads <THREADS>", "-o, --output-dir <OUTPUT_DIR>", "[default: /tmp]",
                 "[default: 4]", "[possible values: kafka, mpi, standalone]", "--seed", "--device", "--terminate-when-clear"):
        assert flag in r.stdout, flag


def test_bad_arguments_are_usage_errors(app):
    assert run(app, "-m", "carrier-pigeon").returncode == 2
    assert run(app, "--no-such-flag").returncode == 2
    assert run(app, "-c").returncode == 2
    r = run(app, "-m", "kafka")
    assert r.returncode == 2 and r.stdout.startswith("Kafka") and "not available" in r.stderr


def test_missing_config_file_is_reported(app, tmp_path):
    r = run(app, "-c", str(tmp_path / "nope.json"), "-o", str(tmp_path))
    assert r.returncode == 1 and r.stdout.startswith("Standalone") and "Failed to read config file" in r.stderr
    r = run(app, "-m", "mpi", "-c", str(tmp_path / "nope.json"), "-o", str(tmp_path))
    assert r.returncode == 1 and "Error while reading config" in r.stderr  # Configuration::read(..).expect(..), main.rs:145


def test_mpi_mode_needs_no_python_and_fails_loudly_without_a_gpu(app, tmp_path):
    """`engine-app -m mpi` forks its own region processes (no Python, no torchrun on the path): with an empty PATH it still
    gets as far as the device check, where it stops -- there is no CPU fallback."""
    import shutil

    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        pytest.skip("a GPU is present: the run itself is covered by the -m gpu tests")
    r = run(app, "-m", "mpi", "-c", os.path.join(GOLDEN, "two_regions_config.json"), "-o", str(tmp_path), env={"PATH": "/nonexistent"})
    assert r.returncode == 1
    assert 1 <= r.stdout.count("MPI") <= 2  # println!("{:?}", args.mode) from the region processes (the second may be stopped before it prints)
    assert "no CUDA device" in r.stderr
    r = run(app, "-m", "mpi", "-c", os.path.join(GOLDEN, "two_regions_config.json"), "-o", str(tmp_path), env={"PATH": "/nonexistent", "RANK": "5", "WORLD_SIZE": "9"})
    assert r.returncode == 1 and "do not fit the 2 regions" in r.stderr


def test_two_region_configuration_parses_and_validates():
    c = Configuration(os.path.join(GOLDEN, "two_regions_config.json"))
    assert c.regions == ["north", "south"] and c.n_regions == 2
    plan = c.travel_plan()
    assert plan["migration"].tolist() == [[0, 30], [20, 0]] and plan["commute"].tolist() == [[0, 40], [25, 0]]
    assert (plan["start_migration_hour"], plan["end_migration_hour"]) == (24, 200)
    cfg = c.engine_config(1)
    assert (cfg.number_of_agents, cfg.grid_size, cfg.hours) == (6000, 200, 240)
    # slots for arrivals: one day's commuters + the migrators of every day in the window (+ slack)
    assert c.arrival_capacity(0) >= 25 + 20 * 8
    assert c.arrival_capacity(1) >= 40 + 30 * 8
    one = c.travel_plan(1)  # fewer ranks than regions: the first regions of the plan
    assert one["n_regions"] == 1 and one["migration"].tolist() == [[0]]


def test_engine_configs_follow_the_region_order(tmp_path):
    """rank r runs travel_plan.regions[r] (MpiTransport::new, mpi_transport.rs:44-52), whatever the order of engine_configs"""
    doc = json.load(open(os.path.join(GOLDEN, "two_regions_config.json")))
    doc["engine_configs"].reverse()
    doc["engine_configs"][0]["config"]["hours"] = 111  # "south" now comes first in engine_configs
    p = tmp_path / "swapped.json"
    p.write_text(json.dumps(doc))
    c = Configuration(str(p))
    assert c.regions == ["north", "south"]
    assert c.engine_config(1).hours == 111 and c.engine_config(0).hours == 240


def test_configuration_errors(tmp_path):
    doc = json.load(open(os.path.join(GOLDEN, "two_regions_config.json")))
    bad = json.loads(json.dumps(doc))
    bad["travel_plan"]["regions"] = ["north", "east"]
    p = tmp_path / "bad.json"
    p.write_text(json.dumps(bad))
    with pytest.raises(EpiError, match="Engine names should match regions"):  # travel_plan_config.rs:60-62
        Configuration(str(p))
    crowded = json.loads(json.dumps(doc))
    crowded["engine_configs"][0]["config"]["geography_parameters"]["grid_size"] = 100  # 100*100 / 6000 < 3
    p.write_text(json.dumps(crowded))
    with pytest.raises(EpiError, match="Not enough space"):  # configuration.rs:111-115
        Configuration(str(p))
    jammed = json.loads(json.dumps(doc))
    jammed["travel_plan"]["commute"]["matrix"] = [[0, 10], [9000, 0]]  # 9000 arrivals > (ceil(200 * 0.2) - 1) * 200 transport cells
    p.write_text(json.dumps(jammed))
    with pytest.raises(EpiError, match="Incoming commuters are more than engine transport capacity"):  # configuration.rs:94-96
        Configuration(str(p))
    disabled = json.loads(json.dumps(doc))
    disabled["travel_plan"]["commute"]["enabled"] = False
    p.write_text(json.dumps(disabled))
    plan = Configuration(str(p)).travel_plan()
    assert plan["commute"] is None and plan["migration"] is not None
    with pytest.raises(EpiError):
        Configuration(str(tmp_path / "missing.json"))


def test_output_writers_use_the_reference_formats(tmp_path):
    rows = np.array([[1, 9, 1, 0, 0, 0, 0], [2, 8, 1, 1, 0, 0, 0]], np.uint32)
    base = write_outputs(str(tmp_path), "north", rows, [(24, 0, 1), (30, 1, 0), (528, 0, 0), (48, 2, 0)], travels=[(24, 1, 3, 1, 0, 0)],
                         region_names=["north", "south"])
    assert os.path.basename(base).startswith("simulation_north_") and os.path.dirname(base).endswith("output")
    assert open(base + ".csv").read() == "hour,susceptible,exposed,infected,hospitalized,recovered,deceased\n1,9,1,0,0,0,0\n2,8,1,1,0,0,0\n"
    assert open(base + "_interventions.json").read() == (
        '[{"hour":24,"intervention":"lockdown","data":{"status":"locked_down"}},{"hour":30,"intervention":"vaccination","data":{}},'
        '{"hour":528,"intervention":"lockdown","data":{"status":"lockdown_revoked"}},{"hour":48,"intervention":"build_new_hospital","data":{}}]')
    assert open(base + "_outgoing_travels.csv").read() == "hr,destination,susceptible,exposed,infected,recovered\n24,south,3,1,0,0\n"
    base2 = write_outputs(str(tmp_path), "0", rows, [])  # standalone: no TravelCounter
    assert not os.path.exists(base2 + "_outgoing_travels.csv") and open(base2 + "_interventions.json").read() == "[]"
