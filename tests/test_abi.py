"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/epi.h declares, parses the
reference's simulation-config JSON, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import json
import os
import re

import pytest

from epirust_b200 import _ffi
from epirust_b200.engine import EpiError, Engine, config_from_json, config_from_json_string, make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def header_symbols():
    text = open(os.path.join(ROOT, "include", "epi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(epi_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _ffi.load()
    syms = header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/epi.h but not exported"
    assert sorted(_ffi.EXPORTS) == syms, "epirust_b200/_ffi.py EXPORTS out of sync with include/epi.h"
    assert b"sm_100a" in L.epi_version()


def test_struct_layout_matches_header():
    assert C.sizeof(_ffi.EpiCounts) == 28
    assert C.sizeof(_ffi.EpiConfig) == 264 + 256  # + population_csv_file[EPI_PATH_MAX]


def test_default_json_parses_like_serde():
    c = config_from_json(os.path.join(GOLDEN, "default_config.json"))
    assert (c.number_of_agents, c.grid_size, c.hours) == (10000, 250, 1080)
    assert (c.public_transport_percentage, c.working_percentage, c.hospital_beds_percentage) == (0.2, 0.7, 0.003)
    assert (c.regular_transmission_start_day, c.high_transmission_start_day, c.last_day) == (5, 6, 26)
    assert (c.regular_transmission_rate, c.high_transmission_rate, c.death_rate) == (0.25, 0.25, 0.035)
    assert (c.exposed_duration, c.pre_symptomatic_duration) == (48, 48)
    assert (c.has_lockdown, c.lockdown_at_number_of_infections, c.essential_workers_population) == (1, 100, 0.1)
    assert c.has_build_new_hospital == 0 and c.n_vaccinations == 0
    assert (c.exposed, c.infected_mild_asymptomatic, c.infected_mild_symptomatic, c.infected_severe) == (1, 0, 0, 0)


def test_auto_pop_fixture_parses():
    # common/src/config/mod.rs:134-189 (should_read_config_with_auto_population)
    c = config_from_json(os.path.join(GOLDEN, "auto_pop_config.json"))
    assert (c.number_of_agents, c.hours, c.grid_size) == (10000, 10000, 250)
    assert c.n_vaccinations == 1 and c.vaccinate_at_hour[0] == 5000 and c.vaccinate_percent[0] == 0.2
    assert (c.infected_mild_asymptomatic, c.infected_mild_symptomatic, c.infected_severe, c.exposed) == (2, 3, 4, 5)
    assert (c.high_transmission_start_day, c.last_day, c.regular_transmission_rate) == (20, 40, 0.025)


def test_starting_infections_default_and_errors():
    base = json.load(open(os.path.join(GOLDEN, "default_config.json")))
    del base["starting_infections"]
    c = config_from_json_string(json.dumps(base))
    assert (c.exposed, c.infected_severe) == (1, 0)  # StartingInfections::default (starting_infections.rs:65-69)
    csv = dict(base)  # Population::Csv (common/src/config/population.rs:30-34, fixture shape of common/config/test/csv_pop.json)
    csv["population"] = {"Csv": {"file": "config/pune_population.csv", "cols": ["age", "sex", "working", "pub_transport"]}}
    c = config_from_json_string(json.dumps(csv))
    assert c.population_csv_file == b"config/pune_population.csv" and c.number_of_agents == 0
    bad = dict(base)
    bad["population"] = {"Csv": {"file": "x.csv"}}  # serde: missing field `cols`
    with pytest.raises(ValueError, match="cols"):
        config_from_json_string(json.dumps(bad))
    bad["population"] = {"Grid": {}}
    with pytest.raises(ValueError, match="unknown variant"):
        config_from_json_string(json.dumps(bad))
    bad = dict(base)
    bad["interventions"] = [{"Curfew": {}}]
    with pytest.raises(ValueError, match="unknown variant"):
        config_from_json_string(json.dumps(bad))
    del bad["disease"]
    with pytest.raises(ValueError):
        config_from_json_string(json.dumps(bad))
    with pytest.raises(ValueError):
        config_from_json_string("{ not json")
    with pytest.raises(ValueError):
        config_from_json("/nonexistent/config.json")


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(EpiError, match="no CPU fallback|CUDA"):
        Engine(make_config(n_agents=100, grid_size=30))


def test_invalid_configs_are_rejected_before_touching_the_device():
    for kw, msg in ((dict(n_agents=0), "number_of_agents"), (dict(grid_size=5), "grid_size"), (dict(working=1.5), "percentage"),
                    (dict(n_agents=100, exposed=200), "starting infections")):
        with pytest.raises(EpiError, match=msg):
            Engine(make_config(**kw))


def test_enable_citizen_state_messages_flag(tmp_path):
    """Config.enable_citizen_state_messages is `#[serde(default)]` bool (common/src/config/mod.rs:54-55): absent = false."""
    import ctypes as C
    L = _ffi.load()
    base = json.load(open(os.path.join(GOLDEN, "default_config.json")))
    on = C.c_int(-1)
    for value, expect in ((None, 0), (False, 0), (True, 1)):
        cfg = dict(base)
        cfg.pop("enable_citizen_state_messages", None)
        if value is not None:
            cfg["enable_citizen_state_messages"] = value
        p = tmp_path / "c.json"
        p.write_text(json.dumps(cfg))
        assert L.epi_config_citizen_state_messages(str(p).encode(), C.byref(on)) == 0
        assert on.value == expect
    assert L.epi_config_citizen_state_messages(str(tmp_path / "missing.json").encode(), C.byref(on)) != 0
