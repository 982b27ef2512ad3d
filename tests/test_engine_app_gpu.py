"""The engine-app command line end to end on the GPU: the files it writes against the oracle's rows for the same seed."""
import csv
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200 import build as B
from epirust_b200 import engine_app as A

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def read_rows(path):
    with open(path) as f:
        rd = list(csv.reader(f))
    assert rd[0] == ["hour", "susceptible", "exposed", "infected", "hospitalized", "recovered", "deceased"]
    return np.array(rd[1:], dtype=np.uint32).reshape(-1, 7)


def test_standalone_cli_writes_the_reference_outputs(tmp_path):
    """BASELINE config #1 (engine/config/default.json values) through the binary."""
    B.build()
    r = subprocess.run([B.APP, "-c", os.path.join(GOLDEN, "default_config.json"), "-o", str(tmp_path), "--seed", "5", "-t", "4"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("Standalone\n") and "Iterations/sec" in r.stdout
    (csv_path,) = glob.glob(str(tmp_path / "output" / "simulation_0_*[0-9].csv"))
    rows_o, events_o, _ = O.oracle_run(O.default_json_config(), seed=5, mode="keyed", threads=4)
    assert (read_rows(csv_path) == rows_o).all()
    (js_path,) = glob.glob(str(tmp_path / "output" / "simulation_0_*_interventions.json"))
    ev = json.load(open(js_path))
    assert [(e["hour"], e["intervention"]) for e in ev] == [(int(h), A.INTERVENTION_NAMES[int(k)]) for h, k, s in events_o]


def test_two_region_cli_one_region_per_gpu(tmp_path):
    """`engine-app -m mpi` with one process and one GPU per region (NCCL all-to-allv) against the multi-region oracle."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cfg_path = os.path.join(GOLDEN, "two_regions_config.json")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "epirust_b200.engine_app", "--launch", "-m", "mpi", "-c", cfg_path, "-o", str(tmp_path), "--seed", "9"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    from epirust_b200.engine import config_from_json_string

    engines, plan = A.read_configuration(cfg_path)
    ocfgs = []
    for e in engines:
        g = config_from_json_string(json.dumps(e["config"]))
        oc = O.EpiConfig()
        for name, _ in O.EpiConfig._fields_:
            v = getattr(g, name)
            if hasattr(v, "__len__"):
                for i in range(len(v)):
                    getattr(oc, name)[i] = v[i]
            else:
                setattr(oc, name, v)
        ocfgs.append(oc)
    orc = O.OracleMultiEngine(ocfgs, seed=9, migration=plan["migration"], commute=plan["commute"], start_migration_hour=plan["start_migration_hour"],
                              end_migration_hour=plan["end_migration_hour"], extra_capacity=2048, threads=2)
    want = np.stack([orc.step(h) for h in range(1, 240)], axis=1)  # [region, hour, 7]
    for k, name in enumerate(plan["regions"]):
        (p,) = glob.glob(str(tmp_path / "output" / f"simulation_{name}_*[0-9].csv"))
        got = read_rows(p)
        assert got.shape == want[k].shape and (got == want[k]).all(), f"region {name}: first differing hour {np.nonzero((got != want[k]).any(axis=1))[0][:3]}"
        (t,) = glob.glob(str(tmp_path / "output" / f"simulation_{name}_*_outgoing_travels.csv"))
        lines = open(t).read().splitlines()
        assert lines[0] == "hr,destination,susceptible,exposed,infected,recovered" and len(lines) > 1
        assert os.path.exists(p[:-4] + "_interventions.json")
