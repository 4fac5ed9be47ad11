"""Known-answer tests of the Go oracle (oracle/go_oracle.py).  The reference's own Go tests assert
types only (rlzero/games/go/test_go_env.py:21-23) and its rule engine (pettingzoo go_base) is not in
the image, so these positions are hand-derived from the rules the reference relies on: parity for Go
is UNPINNED against the reference and pinned only against these vectors."""
import numpy as np
import pytest

from oracle.go_oracle import BLACK, WHITE, GoEnvOracle, IllegalMove, Position, from_flat


def play(pos, moves, n):
    for m in moves:
        pos = pos.play_move(None if m is None else divmod(m, n) if isinstance(m, int) else m)
    return pos


def test_single_stone_capture_and_counts():
    n = 5
    pos = Position(n)
    # black surrounds the white stone at (1,1)
    pos = play(pos, [(0, 1), (1, 1), (1, 0), (4, 4), (2, 1), (4, 3), (1, 2)], n)
    assert pos.board[1, 1] == 0 and pos.caps == (1, 0)
    assert pos.ko is None            # the capturing stone is not surrounded by white: no ko
    assert pos.to_play == WHITE


def test_ko_rule():
    n = 5
    pos = Position(n)
    #   . B W .        black: (0,1) (1,0) (2,1)   white: (0,2) (1,3) (2,2) then white plays (1,1)? build a ko
    pos = play(pos, [(0, 1), (0, 2), (1, 0), (1, 3), (2, 1), (2, 2), (4, 4), (1, 1)], n)
    # white stone at (1,1) is in atari (only liberty (1,2)); black takes at (1,2) capturing (1,1)
    pos = pos.play_move((1, 2))
    assert pos.board[1, 1] == 0 and pos.board[1, 2] == BLACK
    assert pos.ko == (1, 1)
    legal = pos.all_legal_moves()
    assert legal[1 * n + 1] == 0 and legal[n * n] == 1
    with pytest.raises(IllegalMove):
        pos.play_move((1, 1))
    # a move elsewhere lifts the ko
    pos = play(pos, [(4, 0), (3, 4)], n)
    assert pos.ko is None and pos.all_legal_moves()[1 * n + 1] == 1


def test_suicide_is_illegal_but_capture_is_not():
    n = 5
    pos = Position(n)
    # white eye at (0,0): stones (0,1),(1,0); black to move there is suicide
    pos = play(pos, [(4, 4), (0, 1), (4, 3), (1, 0)], n)
    assert pos.to_play == BLACK
    assert not pos.is_move_legal((0, 0))
    assert pos.all_legal_moves()[0] == 0
    # give the two white stones no other liberty: then (0,0) captures and is legal
    pos = play(pos, [(0, 2), (3, 3), (1, 1), (3, 2), (2, 0)], n)      # black 0,2 / 1,1 / 2,0 around them
    assert pos.to_play == WHITE
    pos = pos.play_move(None)
    assert pos.is_move_legal((0, 0))
    pos = pos.play_move((0, 0))
    assert pos.board[0, 1] == 0 and pos.board[1, 0] == 0 and pos.caps[0] == 2


def test_area_scoring_and_komi():
    n = 5
    board = np.zeros((n, n), dtype=np.int8)
    board[:, 1] = BLACK          # black wall: column 0 is black territory
    board[:, 3] = WHITE          # white wall: column 4 white territory, column 2 is dame
    pos = Position(n, komi=0.5, board=board)
    assert pos.score() == (5 + 5) - (5 + 5) - 0.5
    assert pos.result() == -1
    board[2, 2] = BLACK
    assert Position(n, komi=0.5, board=board).score() == 11 - 10 - 0.5
    assert Position(n, komi=0.5, board=board).result() == 1
    # an empty board belongs to nobody
    assert Position(n, komi=7.5).score() == -7.5


def test_env_two_passes_end_the_game_white_wins_by_komi():
    env = GoEnvOracle(5, 7.5)
    env.reset()
    assert list(env.legal_actions()) == list(range(26))
    obs, r, done, info = env.step(25)
    assert not done and r == 0.0 and env.current_player() == 1
    assert obs['observation'].shape == (5, 5, 17) and obs['observation'][:, :, 16].all()
    obs, r, done, info = env.step(25)
    assert done and env.is_terminal() and env.returns() == [-1, 1]
    assert r == -1.0 and isinstance(r, float)            # black's cumulative reward
    assert list(env.legal_actions()) == [25]


def test_env_history_planes_follow_the_mover():
    env = GoEnvOracle(5, 7.5)
    env.reset()
    env.step(0)          # black (0,0)
    env.step(6)          # white (1,1)
    h = env.board_history
    # planes 0,1: (white = last mover, black); planes 2,3: after black's move (black, white)
    assert h[1, 1, 0] and h[0, 0, 1] and h[0, 0, 2] and not h[:, :, 3].any() and not h[:, :, 4:].any()
    obs = env.observe('black_0')
    assert not obs['observation'][:, :, 16].any()
    assert obs['action_mask'][0] == 0 and obs['action_mask'][6] == 0 and obs['action_mask'][25] == 1
    assert env.observe('white_0')['action_mask'].sum() == 0      # not white's turn (go_env.py:161)


def test_from_flat():
    assert from_flat(9, 81) is None and from_flat(9, 10) == (1, 1)


def test_random_games_terminate_and_stay_consistent():
    rs = np.random.RandomState(7)
    for n in (3, 5):
        env = GoEnvOracle(n, 0.5)
        env.reset()
        for t in range(400):
            legal = list(env.legal_actions())
            a = int(legal[rs.randint(len(legal))])
            env.step(a)
            if env.is_terminal():
                break
        b = env._go.board
        # no group without liberties may remain on the board
        from oracle.go_oracle import find_reached
        for r in range(n):
            for c in range(n):
                if b[r, c] != 0:
                    chain, border = find_reached(b, (r, c))
                    assert any(b[p] == 0 for p in border)


def test_one_move_captures_two_groups_and_sets_no_ko():
    """White stones at (1,1) and (1,3) share their last liberty (1,2): black playing there removes both; two captured
    stones never make a ko."""
    n = 5
    pos = play(Position(n), [(0, 1), (1, 1), (2, 1), (1, 3), (1, 0), (4, 4), (0, 3), (4, 3), (2, 3), (4, 2), (1, 4),
                             (4, 1)], n)
    assert pos.to_play == BLACK and pos.board[1, 1] == WHITE and pos.board[1, 3] == WHITE
    pos = pos.play_move((1, 2))
    assert pos.board[1, 1] == 0 and pos.board[1, 3] == 0 and pos.board[1, 2] == BLACK
    assert pos.caps == (2, 0) and pos.ko is None
    # both points are playable for white again (each has empty neighbours now? (1,1): no -- all four neighbours are
    # black, and the black stones around it keep liberties, so it is suicide)
    legal = pos.all_legal_moves()
    assert legal[1 * n + 1] == 0 and legal[1 * n + 3] == 0


def test_corner_capture_and_multi_stone_chain_capture():
    n = 5
    # a corner stone has two liberties
    pos = play(Position(n), [(4, 4), (0, 0), (0, 1), (4, 0), (1, 0)], n)
    assert pos.board[0, 0] == 0 and pos.caps == (1, 0)
    assert pos.ko is None            # (1,0) is not enclosed by white stones
    # a two-stone chain on the edge: white (0,1),(0,2); black needs (0,0),(1,1),(1,2),(0,3)
    pos = play(Position(n), [(0, 0), (0, 1), (1, 1), (0, 2), (1, 2), (4, 4), (0, 3)], n)
    assert pos.board[0, 1] == 0 and pos.board[0, 2] == 0 and pos.caps == (2, 0) and pos.ko is None


def test_suicide_of_a_chain_is_illegal():
    """Black (0,0) + a stone at (0,1) would form a two-stone chain with no liberty and capture nothing."""
    n = 5
    pos = play(Position(n), [(0, 0), (1, 0), (4, 4), (1, 1), (4, 3), (0, 2)], n)
    assert pos.to_play == BLACK
    assert not pos.is_move_legal((0, 1)) and pos.all_legal_moves()[1] == 0
    with pytest.raises(IllegalMove):
        pos.play_move((0, 1))
    # once the white stone at (0,2) is itself in atari on that point, the same move captures it and is legal:
    # black takes (0,3) and (1,2), white passes
    pos = play(pos, [(0, 3), None, (1, 2), None], n)
    assert pos.to_play == BLACK and pos.is_move_legal((0, 1))
    pos = pos.play_move((0, 1))
    assert pos.board[0, 2] == 0 and pos.caps[0] == 1 and pos.board[0, 0] == BLACK and pos.board[0, 1] == BLACK


def test_tromp_taylor_counts_only_single_coloured_regions():
    """5x5: a black wall on column 1 and a white wall on column 3: column 0 is black territory, column 4 white
    territory, column 2 touches both (dame).  Black 5 + 5, white 5 + 5, komi decides."""
    n = 5
    moves = []
    for r in range(n):
        moves += [(r, 1), (r, 3)]
    pos = play(Position(n, komi=0.5), moves, n)
    assert pos.score() == -0.5 and pos.result() == -1
    pos2 = play(Position(n, komi=0.0), moves, n)
    assert pos2.score() == 0.0 and pos2.result() == 0
    # a black stone inside the white territory that white does not bother to capture makes column 4 dame too
    pos3 = play(Position(n, komi=0.5), moves + [(2, 4), None], n)
    assert pos3.score() == 10 + 1 - 5 - 0.5
