"""GPU edge cases of the search path: empty batches, pools and tables that run out (device fault bits -> host
exceptions, no memory damage), odd batch sizes."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_empty_batches_are_no_ops():
    from rlzero_b200 import _lib as L
    lib = L.load()
    g = L.GameDesc(15, 5, 225, 256)
    s = L.stream_ptr()
    dummy = torch.zeros(64, dtype=torch.int32, device='cuda')
    assert lib.rz_gomoku_reset(C.byref(g), L.ptr(dummy), L.ptr(dummy), 0, 0, s) == 0
    assert lib.rz_gomoku_step(C.byref(g), L.ptr(dummy), L.ptr(dummy), L.ptr(dummy), None, None, 0, s) == 0
    assert lib.rz_net_conv3x3_tc3(L.ptr(dummy), L.ptr(dummy), L.ptr(dummy), None, L.ptr(dummy[32:]), 0, 15, 15, 16, 1, 0,
                                  s) == 0
    hd = L.HeadsDesc()
    hd.board_size, hd.action_stride = 15, 256
    for name in ('w1x1', 'b1x1', 'wp', 'bp', 'wv1', 'bv1', 'wv2', 'bv2', 'wtc_hi', 'wtc_lo'):
        setattr(hd, name, dummy.data_ptr())
    assert lib.rz_net_heads_tc(C.byref(hd), L.ptr(dummy), L.ptr(dummy), L.ptr(dummy), 0, s) == 0
    assert lib.rz_net_head_features(C.byref(hd), L.ptr(dummy), L.ptr(dummy), 0, s) == 0
    torch.cuda.synchronize()
    assert int(dummy.abs().sum()) == 0


def test_pool_overflow_sets_a_fault_and_keeps_searching():
    """More playouts than expanded-node capacity: the device marks the leaf RZ_CHILD_OVERFLOW, sets the fault bit,
    and later visits re-evaluate it like a leaf; the host raises, nothing is written out of bounds."""
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    f = SearchForest(3, 6, 4, n_playout=60, max_nodes=10, max_carry=0)
    guard = f.edge_N.clone()
    f.search(ClosedFormEvaluator(EVAL_HASH))
    bits = f.faults()
    assert all(int(b) & L.FAULT_POOL_OVERFLOW for b in bits)
    assert f.n_nodes.cpu().tolist() == [10, 10, 10]
    assert f.root_N.cpu().tolist() == [60, 60, 60]          # every playout was still backed up
    with pytest.raises(RuntimeError):
        f.raise_faults()
    assert guard.shape == f.edge_N.shape
    # leaf-parallel mode overflows the same way
    f2 = SearchForest(2, 6, 4, n_playout=60, max_nodes=10, max_carry=0, leaves_per_tree=8)
    f2.search(ClosedFormEvaluator(EVAL_HASH))
    assert all(int(b) & L.FAULT_POOL_OVERFLOW for b in f2.faults())
    assert f2.root_N.cpu().tolist() == [60, 60] and f2.n_nodes.cpu().tolist() == [10, 10]


def test_ln_table_overflow_is_reported():
    from oracle.evaluators import EVAL_KAT
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    f = SearchForest(1, 3, 3, n_playout=40, ln_table_len=16)
    f.search(ClosedFormEvaluator(EVAL_KAT))
    assert int(f.faults()[0]) & L.FAULT_LN_TABLE


@pytest.mark.parametrize('G', [1, 5, 13, 131])
def test_odd_batch_sizes_agree_with_each_other(G):
    """Trees are independent: tree g of a G-tree forest equals tree 0 of a 1-tree forest on the same position
    (grids are rounded up to 4 warps per block; the tail warps must neither run nor be skipped)."""
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    rs = np.random.RandomState(G)
    moves = [[int(a) for a in rs.permutation(36)[:g % 7]] for g in range(G)]
    f = SearchForest(G, 6, 4, n_playout=40)
    f.set_positions(moves)
    f.search(ClosedFormEvaluator(EVAL_HASH))
    f.raise_faults()
    visits = f.root_stats()[0]
    for g in sorted(set([0, G // 2, G - 1])):
        f1 = SearchForest(1, 6, 4, n_playout=40)
        f1.set_positions([moves[g]])
        f1.search(ClosedFormEvaluator(EVAL_HASH))
        assert np.array_equal(f1.root_stats()[0][0], visits[g]), g
