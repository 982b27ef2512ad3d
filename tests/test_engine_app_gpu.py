"""The engine-app command line end to end on the GPU: the files it writes against the oracle's rows for the same seed."""
import csv
import glob
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200 import build as B
from epirust_b200.engine import Configuration, device_count

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
INTERVENTION_NAMES = ("lockdown", "vaccination", "build_new_hospital")  # lockdown.rs:104-114, vaccination.rs:58-64, hospital.rs:78-84


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
    assert [(e["hour"], e["intervention"]) for e in ev] == [(int(h), INTERVENTION_NAMES[int(k)]) for h, k, s in events_o]


def test_citizen_state_messages_stream(tmp_path):
    """Config.enable_citizen_state_messages (common/src/config/mod.rs:54-55): one CitizenStatesAtHr JSON line per simulated hour
    (listeners/events_kafka_producer.rs:62-100, models/events/citizen_state.rs:26-62), the stand-in for the
    `citizen_states_updated` topic.  Every line must agree with the oracle's agents of that hour: state letter and location per slot."""
    B.build()
    cfg = json.load(open(os.path.join(GOLDEN, "default_config.json")))
    cfg["population"]["Auto"]["number_of_agents"] = 600
    cfg["geography_parameters"]["grid_size"] = 80
    cfg["hours"] = 60
    cfg["starting_infections"] = {"infected_mild_asymptomatic": 5, "infected_mild_symptomatic": 5, "infected_severe": 5, "exposed": 20}
    cfg["enable_citizen_state_messages"] = True
    path = tmp_path / "with_states.json"
    path.write_text(json.dumps(cfg))
    r = subprocess.run([B.APP, "-c", str(path), "-o", str(tmp_path), "--seed", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    (csv_path,) = glob.glob(str(tmp_path / "output" / "simulation_0_*[0-9].csv"))
    rows = read_rows(csv_path)
    (st_path,) = glob.glob(str(tmp_path / "output" / "simulation_0_*_citizen_states.jsonl"))
    lines = open(st_path).read().splitlines()
    assert json.loads(lines[-1]) == {"simulation_ended": True}
    assert len(lines) == len(rows) + 1
    from epirust_b200.engine import config_from_json
    gcfg = config_from_json(str(path))
    ocfg = O.EpiConfig()
    for name, _ in O.EpiConfig._fields_:
        v = getattr(gcfg, name)
        if hasattr(v, "__len__") and not isinstance(v, (bytes, str)):
            for i in range(len(v)):
                getattr(ocfg, name)[i] = v[i]
        else:
            setattr(ocfg, name, v)
    orc = O.OracleEngine(ocfg, seed=3)
    letters = "seird"
    for k, line in enumerate(lines[:-1]):
        msg = json.loads(line)
        hour = k + 1
        assert msg["hr"] == hour
        orc.step(hour)
        st = orc.get_state()
        got = msg["citizen_states"]
        assert len(got) == 600
        for slot, c in enumerate(got):
            assert c["citizen_id"] == "00000000-0000-4000-8000-%012x" % slot
            assert c["state"] == letters[int(st["st"][slot]) & 7]
            assert (c["location"]["x"], c["location"]["y"]) == (int(st["cell_x"][slot]), int(st["cell_y"][slot]))
        counts = [sum(1 for c in got if c["state"] == ch) for ch in letters]
        assert counts == [int(rows[k][1]), int(rows[k][2]), int(rows[k][3]) + int(rows[k][4]), int(rows[k][5]), int(rows[k][6])]
    # the flag is off by default: no such file
    cfg["enable_citizen_state_messages"] = False
    path.write_text(json.dumps(cfg))
    r2 = subprocess.run([B.APP, "-c", str(path), "-o", str(tmp_path / "plain"), "--seed", "3"], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr
    assert not glob.glob(str(tmp_path / "plain" / "output" / "*_citizen_states.jsonl"))


def oracle_of(c, seed, hours):
    ocfgs = []
    for r in range(c.n_regions):
        g = c.engine_config(r)
        oc = O.EpiConfig()
        for name, _ in O.EpiConfig._fields_:
            v = getattr(g, name)
            if hasattr(v, "__len__") and not isinstance(v, (bytes, str)):
                for i in range(len(v)):
                    getattr(oc, name)[i] = v[i]
            else:
                setattr(oc, name, v)
        ocfgs.append(oc)
    plan = c.travel_plan()
    extra = max(c.arrival_capacity(r) for r in range(c.n_regions))
    orc = O.OracleMultiEngine(ocfgs, seed=seed, migration=plan["migration"], commute=plan["commute"], start_migration_hour=plan["start_migration_hour"],
                              end_migration_hour=plan["end_migration_hour"], extra_capacity=extra, threads=2)
    return np.stack([orc.step(h) for h in range(1, hours)], axis=1)  # [region, hour, 7]


@pytest.mark.parametrize("data_plane", ["peer_memory", "nccl"])
def test_two_region_cli_one_region_per_gpu(tmp_path, data_plane):
    """`engine-app -m mpi`: the binary forks one process per region, one region per GPU, the traveller exchange between them --
    with nothing but the binary on PATH (no Python, no torchrun) -- against the multi-region oracle.  Both data planes: the fused
    kernel over NVLink peer memory (default) and leave -> grouped ncclSend / ncclRecv -> arrive (EPI_NO_PEER=1, what a box whose
    GPUs cannot map each other's memory falls back to)."""
    if device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    B.build()
    cfg_path = os.path.join(GOLDEN, "two_regions_config.json")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "PYTHONPATH", "EPI_NO_PEER")}
    env["PATH"] = "/nonexistent"
    if data_plane == "nccl":
        env["EPI_NO_PEER"] = "1"
    r = subprocess.run([B.APP, "-m", "mpi", "-c", cfg_path, "-o", str(tmp_path), "--seed", "9"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("MPI\n") == 2
    c = Configuration(cfg_path)
    want = oracle_of(c, 9, 240)
    for k, name in enumerate(c.regions):
        (p,) = glob.glob(str(tmp_path / "output" / f"simulation_{name}_*[0-9].csv"))
        got = read_rows(p)
        assert got.shape == want[k].shape and (got == want[k]).all(), f"region {name}: first differing hour {np.nonzero((got != want[k]).any(axis=1))[0][:3]}"
        (t,) = glob.glob(str(tmp_path / "output" / f"simulation_{name}_*_outgoing_travels.csv"))
        lines = open(t).read().splitlines()
        assert lines[0] == "hr,destination,susceptible,exposed,infected,recovered" and len(lines) > 1
        assert os.path.exists(p[:-4] + "_interventions.json")


def test_two_ranks_through_the_c_abi_only():
    """Two region processes that call nothing but epi_* (ctypes): epi_comm_unique_id -> epi_run_region on each rank."""
    if device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import multiprocessing as mp

    cfg_path = os.path.join(GOLDEN, "two_regions_config.json")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    from epirust_b200.engine import comm_unique_id

    uid = comm_unique_id()
    procs = [ctx.Process(target=_rank_main, args=(cfg_path, r, 2, uid, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = oracle_of(Configuration(cfg_path), 9, 240)
    for r in range(2):
        assert got[r].shape == want[r].shape and (got[r] == want[r]).all()


def _rank_main(cfg_path, rank, world, uid, seed, q):
    from epirust_b200.engine import Configuration as Cfg

    rows, _ = Cfg(cfg_path).run_region(rank, world, uid, seed=seed, device=rank)
    q.put((rank, rows))
