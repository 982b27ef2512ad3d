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


def check_numbering(s, n, H, n_off):
    """The population is the reference's (citizen c -> house c % H, office c % O, grid.rs:108-113) whatever the numbering;
    the default numbering is house by house (DESIGN.md "Agent numbering")."""
    home, work = s["home"].astype(np.int64), s["work"].astype(np.int64)
    assert (np.bincount(home, minlength=H) == np.bincount(np.arange(n) % H, minlength=H)).all()
    ws = (s["st"] >> 13) & 3
    ok = np.zeros(n, bool)
    for k in range(4):  # at most HOME_SIZE^2 = 4 citizens per house
        c = home + k * H
        ok |= (c < n) & (work == c % n_off)
    assert ok[ws != 3].all()
    assert (np.diff(home) >= 0).all()  # house by house

@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("seed", [1, 12345678901234567])
def test_population_factory_matches_oracle(kw, seed):
    ours = build_population(make_config(**kw), seed=seed)
    orc = O.OracleEngine(O.make_config(**kw), seed=seed).get_state()
    for f in STATE_FIELDS:
        assert (ours[f] == orc[f]).all(), f"{f} differs"


def test_numbering_is_a_relabelling_of_the_reference_rule():
    # grid.rs:108-113: creation number c -> house c % H, office c % O; at most HOME_SIZE^2 = 4 per house
    n = 20000
    s = build_population(make_config(n_agents=n, grid_size=250), seed=3)
    check_numbering(s, n, 6250, 125)  # geography for G = 250
    assert len(set(zip(s["cell_x"].tolist(), s["cell_y"].tolist()))) == n  # distinct start cells
