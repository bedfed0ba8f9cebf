"""Design-of-experiments batching (reference: mainDoE.py): every (run, segment) row of the batch
equals the oracle run on that run's tables, bit for bit on the CPU test double."""
import copy

import numpy as np
import pytest

from hspsquared_b200 import doe as DOE
from oracle import cases as CASES

from . import parity

DOE_LIST = [  # the reference's own docstring example, PERLND lines (mainDoE.py:27-57)
    [1, "PERLND/PWATER/PARAMETERS", "P001", "FOREST", 0.02],
    [1, "PERLND/PWATER/PARAMETERS", "P001", "INFILT", 0.15],
    [2, "PERLND/PWATER/PARAMETERS", "P001", "FOREST", 0.01],
    [2, "PERLND/PWATER/PARAMETERS", "P001", "INFILT", 0.20],
    [5, "PERLND/PWATER/PARAMETERS", "P001", "INFILT", 0.05],
    [10, "PERLND/SNOW/PARAMETERS", "P001", "MWATER", 0.09],
    [10, "IMPLND/SNOW/PARAMETERS", "I001", "MWATER", 0.09],
    [12, "PERLND/PWATER/MONTHLY/CEPSC", "P001", "MAR", 0.04],
    [13, "PERLND/PWATER/MONTHLY/CEPSC", "P002", "MAY", 0.05],
    [14, "PERLND/SNOW/FLAGS", "P001", "ICEFG", 0],
    [15, "PERLND/SNOW/FLAGS", "P001", "ICEFG", 1],
]


@pytest.fixture(scope="module", autouse=True)
def _use_emu(emu_library):
    yield


def test_make_runlist_matches_the_reference_parsing():
    rl = DOE.make_runlist(DOE_LIST)
    assert list(rl) == ["1", "2", "5", "10", "12", "13", "14", "15"]
    assert rl["1"][("PERLND", "PWATER", "P001")]["PARAMETERS"] == {"FOREST": 0.02, "INFILT": 0.15}
    assert rl["12"][("PERLND", "PWATER", "P001")]["MONTHLY_CEPSC"] == {"MAR": 0.04}
    assert isinstance(rl["14"][("PERLND", "SNOW", "P001")]["FLAGS"]["ICEFG"], float)
    assert ("IMPLND", "SNOW", "I001") in rl["10"]


def test_doe_batch_equals_oracle_run_by_run():
    parity.check_doe(DOE_LIST)
