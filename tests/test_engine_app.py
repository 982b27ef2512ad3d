"""The outer boundary: the engine-app command line (engine-app/src/main.rs:60-177) and the multi-region Configuration
(common/src/config/configuration.rs:221-315).  CPU-only checks here; the runs themselves are in the -m gpu tests."""
import json
import os
import subprocess

import numpy as np
import pytest

from epirust_b200 import build as B
from epirust_b200 import engine_app as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def app():
    B.build()
    assert os.path.exists(B.APP)
    return B.APP


def run(app, *args):
    return subprocess.run([app, *args], capture_output=True, text=True, timeout=120)


def test_help_lists_the_reference_flags(app):
    r = run(app, "--help")
    assert r.returncode == 0
    for flag in ("-c, --config <FILE>", "-m, --mode <MODE>", "-i, --id <ID>", "-t, --threads <THREADS>", "-o, --output-dir <OUTPUT_DIR>", "[default: /tmp]",
                 "[default: 4]", "[possible values: kafka, mpi, standalone]", "--seed", "--device"):
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


def test_two_region_configuration_parses_and_validates():
    engines, plan = A.read_configuration(os.path.join(GOLDEN, "two_regions_config.json"))
    assert [e["engine_id"] for e in engines] == ["north", "south"] and plan["regions"] == ["north", "south"]
    assert plan["migration"].tolist() == [[0, 30], [20, 0]] and plan["commute"].tolist() == [[0, 40], [25, 0]]
    assert (plan["start_migration_hour"], plan["end_migration_hour"]) == (24, 200)
    A.validate_configuration(engines, plan)
    # slots for arrivals: one day's commuters + the migrators of every day in the window (+ slack)
    assert A.arrival_capacity(plan, 0, 240) >= 25 + 20 * 8
    assert A.arrival_capacity(plan, 1, 240) >= 40 + 30 * 8


def test_configuration_errors(tmp_path):
    doc = json.load(open(os.path.join(GOLDEN, "two_regions_config.json")))
    bad = json.loads(json.dumps(doc))
    bad["travel_plan"]["regions"] = ["north", "east"]
    p = tmp_path / "bad.json"
    p.write_text(json.dumps(bad))
    with pytest.raises(A.ConfigError, match="Engine names should match regions"):  # configuration.rs:243-245
        A.read_configuration(str(p))
    crowded = json.loads(json.dumps(doc))
    crowded["engine_configs"][0]["config"]["geography_parameters"]["grid_size"] = 100  # 100*100 / 6000 < 3
    p.write_text(json.dumps(crowded))
    engines, plan = A.read_configuration(str(p))
    with pytest.raises(A.ConfigError, match="Not enough space"):  # configuration.rs:299-303
        A.validate_configuration(engines, plan)
    disabled = json.loads(json.dumps(doc))
    disabled["travel_plan"]["commute"]["enabled"] = False
    p.write_text(json.dumps(disabled))
    _, plan = A.read_configuration(str(p))
    assert plan["commute"] is None and plan["migration"] is not None


def test_output_writers_use_the_reference_formats(tmp_path):
    rows = np.array([[1, 9, 1, 0, 0, 0, 0], [2, 8, 1, 1, 0, 0, 0]], np.uint32)
    base = A.output_file_format(str(tmp_path), "north")
    assert os.path.basename(base).startswith("simulation_north_") and os.path.dirname(base).endswith("output")
    A.write_outputs(base, rows, [(24, 0, 1), (30, 1, 0), (528, 0, 0), (48, 2, 0)], travels=[(24, "south", 3, 1, 0, 0)])
    assert open(base + ".csv").read() == "hour,susceptible,exposed,infected,hospitalized,recovered,deceased\n1,9,1,0,0,0,0\n2,8,1,1,0,0,0\n"
    assert open(base + "_interventions.json").read() == (
        '[{"hour":24,"intervention":"lockdown","data":{"status":"locked_down"}},{"hour":30,"intervention":"vaccination","data":{}},'
        '{"hour":528,"intervention":"lockdown","data":{"status":"lockdown_revoked"}},{"hour":48,"intervention":"build_new_hospital","data":{}}]')
    assert open(base + "_outgoing_travels.csv").read() == "hr,destination,susceptible,exposed,infected,recovered\n24,south,3,1,0,0\n"
