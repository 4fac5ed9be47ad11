"""CPU tests: the plain-C Go restatement (oracle/c/rz_go_oracle.c) against the Python one (oracle/go_oracle.py,
MiniGo's published rules at the reference's call sites, go_env.py:98-112,168-210) -- rules on random games to the
end, and whole searches bit for bit.  Parity with the reference itself stays UNPINNED for Go (its engine,
pettingzoo's go_base, is absent); these tests make the fast oracle say what the readable one says."""
import numpy as np
import pytest

from oracle import build_oracle, pyoracle
from oracle.evaluators import EVAL_HASH, EVAL_KAT, make_policy_value_fn
from oracle.go_oracle import GoSearchBoard, Position, from_flat


@pytest.mark.parametrize('n,komi,plies,cap', [(3, 0.5, 40, 0), (5, 2.5, 120, 0), (9, 7.5, 250, 0), (9, 7.5, 400, 60),
                                              (19, 7.5, 120, 0)])
def test_rules_on_random_games(n, komi, plies, cap):
    G = 24 if n < 19 else 6
    out = build_oracle.go_random_games(G, n, [plies * (g + 1) // G for g in range(G)], komi, cap, seed=n)
    captures = 0
    for g in range(G):
        pos = Position(n, komi)
        for a in out['moves'][g][:out['played'][g]]:
            before = int(np.count_nonzero(pos.board))
            pos = pos.play_move(from_flat(n, int(a)))          # raises IllegalMove if the C side played an illegal move
            captures += int(np.count_nonzero(pos.board)) < before
        assert np.array_equal(pos.board.reshape(-1), out['cell'][g])
        assert (pos.ko[0] * n + pos.ko[1] if pos.ko is not None else -1) == out['ko'][g]
        assert (0 if pos.to_play == 1 else 1) == out['to_play'][g]
        over = pos.is_game_over() or (cap > 0 and pos.n >= cap)
        assert bool(out['over'][g]) == over
        assert pos.score() == out['score'][g]
        if not over:
            assert np.array_equal(pos.all_legal_moves(), out['legal'][g])
    assert captures > 0 or n == 19


@pytest.mark.parametrize('n,n_playout,plies,eval_id,rule,cap', [
    (3, 80, 6, EVAL_HASH, 0, 0), (5, 150, 20, EVAL_HASH, 0, 0), (5, 150, 30, EVAL_HASH, 1, 0), (5, 60, 20, EVAL_KAT, 0, 0),
    (9, 120, 60, EVAL_HASH, 0, 0), (5, 80, 4, EVAL_HASH, 0, 8), (9, 60, 140, EVAL_HASH, 1, 0)])
def test_search_matches_the_python_oracle(n, n_playout, plies, eval_id, rule, cap):
    G, komi = 5, 2.5
    games = build_oracle.go_random_games(G, n, [plies * (g + 1) // G for g in range(G)], komi, cap, seed=7 * n + rule)
    lists = [games['moves'][g][:games['played'][g]].tolist() for g in range(G) if not games['over'][g]]
    assert lists
    visits, w, rn, rw = build_oracle.go_search_batch(n, lists, n_playout, komi, cap, 5.0, rule, eval_id)
    A = n * n + 1
    for g, moves in enumerate(lists):
        b = GoSearchBoard(n, komi, cap)
        for a in moves:
            b.step(a)
        s = pyoracle.Search(make_policy_value_fn(eval_id), n_playout, 5, rule=rule)
        s.simulate(b, 1.0)
        assert np.array_equal(visits[g], s.root_visits(A)), g
        assert np.array_equal(w[g], s.root_values(A)), g
        assert rn[g] == s.root.n and rw[g] == s.root.w
