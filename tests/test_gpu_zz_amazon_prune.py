"""Line search by bounds against the exhaustive search on BASELINE config 3 (com-amazon K=200): the largest case of
tests/test_gpu_prune.py, collected last (file name) and under a hard timeout — the round's last GPU call was killed at its
budget limit inside this case on a slow box; whatever happens here, the other results are on the screen."""
import pytest

from test_gpu_prune import _baseline_graph_case

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600, method="thread")
def test_bounds_change_nothing_on_com_amazon(graphs):
    """8 steps from the synthetic F0 of the bench: far enough for a good part of the nodes to have stopped moving (that is
    where the bounds exclude the most); accepted step of every node, rows, sumF and LLH identical in both engines."""
    asked, searched = _baseline_graph_case(graphs, "com-amazon", 200, 8)
    assert searched < 0.75 * asked, f"{searched} of {asked}"
