"""Native BIF reader against the reference's own reader (through the exported fixtures) and its documented layout."""
import os

import numpy as np
import pytest

import pgmpy_b200 as px
from pgmpy_b200.readwrite import BIFReader

BIF = """
network unknown {
}
variable rain {  // a comment
  type discrete [ 2 ] { yes, no };
}
variable sprinkler {
  type discrete [ 2 ] { on, off };
}
variable grass {
  type discrete [ 3 ] { dry, damp, wet };
}
probability ( rain ) {
  table 0.2, 0.8;
}
probability ( sprinkler | rain ) {
  (yes) 0.01, 0.99;
  (no) 0.4, 0.6;
}
probability ( grass | sprinkler, rain ) {
  (off, no) 1.0, 0.0, 0.0;
  (on, yes) 0.0, 0.01, 0.99;
  (on, no) 0.05, 0.15, 0.8;
  (off, yes) 0.1, 0.2, 0.7;
}
"""


def test_bif_string_layout():
    m = BIFReader(string=BIF).get_model()
    assert m.nodes() == ["rain", "sprinkler", "grass"]
    assert sorted(m.edges()) == [("rain", "grass"), ("rain", "sprinkler"), ("sprinkler", "grass")]
    g = m.get_cpds("grass")
    assert g.variables == ["grass", "sprinkler", "rain"] and list(g.cardinality) == [3, 2, 2]
    # columns follow itertools.product(sprinkler states, rain states): (on,yes) (on,no) (off,yes) (off,no)
    np.testing.assert_array_equal(g.get_values(), [[0.0, 0.05, 0.1, 1.0], [0.01, 0.15, 0.2, 0.0], [0.99, 0.8, 0.7, 0.0]])
    assert g.state_names["grass"] == ["dry", "damp", "wet"]
    np.testing.assert_array_equal(m.get_cpds("sprinkler").get_values(), [[0.01, 0.4], [0.99, 0.6]])
    assert m.check_model()
    assert [c.variable for c in m.get_cpds()] == ["grass", "rain", "sprinkler"]  # sorted-name order, like the reference


@pytest.mark.parametrize("name", ["asia", "alarm", "hepar2", "win95pts", "pathfinder", "munin"])
def test_bif_reader_matches_reference_reader(name):
    """Same model as the fixture that was exported through pgmpy's own BIFReader (oracle/export_models.py)."""
    path = f"/root/reference/pgmpy/utils/example_models/{name}.bif.gz"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    got = BIFReader(path).get_model()
    want = px.get_example_model(name)
    assert got.nodes() == want.nodes()
    assert sorted(got.edges()) == sorted(want.edges())
    for v in want.nodes():
        a, b = got.get_cpds(v), want.get_cpds(v)
        assert a.variables == b.variables
        assert a.state_names == b.state_names
        np.testing.assert_array_equal(a.values, b.values)
