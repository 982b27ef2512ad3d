"""CPU check of host logic: the product's Auto population factory (csrc/host_model.cpp, via the host-only C-ABI entry
epi_build_population) against the oracle's independent restatement of Grid::generate_population + citizen_factory
(engine/src/geography/grid.rs:83-155, citizen/citizen_factory.rs:31-134).  Same Philox-keyed init draws -> identical."""
import numpy as np
import pytest

import oracle_ffi as O
from epirust_b200.engine import STATE_FIELDS, build_population, make_config

CASES = [
    dict(n_agents=10000, grid_size=250, exposed=1),                                  # default.json: 6250 houses, 1.6 agents / house
    dict(n_agents=3000, grid_size=250, exposed=7, asym=2, mild=3, severe=4),          # fewer agents than houses
    dict(n_agents=5000, grid_size=120, exposed=5, lockdown=(100, 0.25)),             # 3.5 agents / house, essential workers drawn
    dict(n_agents=999, grid_size=57, exposed=3),                                      # ragged: odd strip widths
]


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("seed", [1, 12345678901234567])
def test_population_factory_matches_oracle(kw, seed):
    ours = build_population(make_config(**kw), seed=seed)
    orc = O.OracleEngine(O.make_config(**kw), seed=seed).get_state()
    for f in STATE_FIELDS:
        assert (ours[f] == orc[f]).all(), f"{f} differs"


def test_house_by_house_numbering_is_a_relabelling_of_the_reference_rule():
    # grid.rs:108-113: creation number c -> house c % H, office c % O; at most HOME_SIZE^2 = 4 per house
    kw = dict(n_agents=20000, grid_size=250)
    s = build_population(make_config(**kw), seed=3)
    H, n_off = 6250, 125  # geography for G = 250
    home = s["home"].astype(np.int64)
    assert (np.diff(home) >= 0).all()
    rank = np.arange(len(home)) - np.searchsorted(home, home, side="left")
    creation = home + rank * H
    assert sorted(creation.tolist()) == list(range(len(home)))
    ws = (s["st"] >> 13) & 3
    assert (s["work"][ws != 3] == (creation % n_off)[ws != 3]).all()
    assert np.bincount(home, minlength=H).max() <= 4
    assert len(set(zip(s["cell_x"].tolist(), s["cell_y"].tolist()))) == len(home)  # distinct start cells
