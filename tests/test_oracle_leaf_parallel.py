"""CPU tests of the oracle's restatement of the leaf-parallel wave (oracle.pyoracle.Search.wave): the virtual
statistics are taken off exactly, the playout budget is met exactly, and with one leaf per wave the wave IS the
reference's sequential playout (bit-identical statistics)."""
import numpy as np
import pytest

from oracle import pyoracle
from oracle.evaluators import EVAL_HASH, make_policy_value_fn


def _walk(node):
    yield node
    for ch in node.children.values():
        yield from _walk(ch)


@pytest.mark.parametrize('size,k,n_playout,K', [(3, 3, 50, 4), (6, 4, 120, 8), (5, 4, 200, 16)])
def test_wave_restores_virtual_statistics_and_meets_the_budget(size, k, n_playout, K):
    fn = make_policy_value_fn(EVAL_HASH)
    b = pyoracle.Board(size, k)
    b.reset()
    b.step(size + 1)
    s = pyoracle.Search(fn, n_playout, 2.5, leaves_per_wave=K, virtual_loss=1.0)
    acts, probs = s.simulate(b, 1.0)
    assert s.root.n == n_playout and abs(probs.sum() - 1.0) < 1e-12
    assert sum(ch.n for ch in s.root.children.values()) == n_playout - 1
    for nd in _walk(s.root):
        assert nd.n >= 0
        # a node's visits = its own evaluations (leaf visits) + its children's visits
        if nd.children:
            assert nd.n >= sum(ch.n for ch in nd.children.values())
        assert abs(nd.w) <= nd.n + 1e-9          # |values| <= 1: no virtual loss left behind


def test_single_leaf_wave_is_the_sequential_playout():
    fn = make_policy_value_fn(EVAL_HASH)
    b = pyoracle.Board(6, 4)
    b.reset()
    seq = pyoracle.Search(fn, 90, 5.0)
    seq.simulate(b, 1.0)
    one = pyoracle.Search(fn, 90, 5.0, leaves_per_wave=1)
    for _ in range(90):
        one.wave(b, 1)
    assert np.array_equal(seq.root_visits(36), one.root_visits(36))
    assert [float(x).hex() for x in seq.root_values(36)] == [float(x).hex() for x in one.root_values(36)]
    assert float(seq.root.w).hex() == float(one.root.w).hex()
