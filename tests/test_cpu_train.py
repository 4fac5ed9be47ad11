"""CPU tests of the training-side host logic: TrainPipeline.get_equi_data against the fixture the
live reference produced (scripts/make_golden_next.py)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EQUI = json.load(open(os.path.join(HERE, 'golden', 'equi.json')))['cases']


@pytest.mark.parametrize('case', EQUI, ids=lambda c: 'equi_%d' % c['size'])
def test_host_get_equi_data_matches_reference(case):
    """TrainPipeline.get_equi_data (host, numpy) reproduces the live reference's output bit for bit."""
    from rlzero_b200.train_pipeline import TrainPipeline
    size, n = case['size'], case['n']
    rs = np.random.RandomState(case['seed'])
    states = rs.randint(0, 2, size=(n, 4, size, size)).astype(np.float64)
    pis = rs.rand(n, size * size)
    zs = rs.choice([-1.0, 0.0, 1.0], size=n)
    tp = TrainPipeline.__new__(TrainPipeline)
    tp.board_size = size
    out = tp.get_equi_data(list(zip(states, pis, zs)))
    assert len(out) == 8 * n
    for o, s_ref, p_ref, z_ref in zip(out, case['states'], case['pis'], case['zs']):
        assert o[0].astype(np.int8).reshape(-1).tolist() == s_ref
        assert [float(x).hex() for x in o[1]] == p_ref
        assert float(o[2]) == z_ref


