"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Go path.  PARITY UNPINNED.

The reference's ``GoEnv`` (``rlzero/games/go/go_env.py``) is a thin wrapper over
``pettingzoo.classic.go.{go_base, coords}``: a third-party dependency that is NOT vendored under
``/root/reference``, NOT installed in this image and NOT version-pinned by the reference (no
requirements file).  ``go_base`` is PettingZoo's copy of MiniGo's ``go.py``; this module restates
that published algorithm (Tromp-Taylor rules: capture of liberty-less opponent groups, no suicide,
simple ko, pass, game over after two consecutive passes, area scoring minus komi) at exactly the
call sites the reference uses:

  go_env.py:172      Position.play_move(coords.from_flat(action))      -> ``Position.play_move``
  go_env.py:185,190  Position.is_game_over(), Position.result()        -> ``is_game_over``, ``result``
  go_env.py:193-194  Position.all_legal_moves()                        -> ``all_legal_moves``
  go_env.py:213      go_base.Position(board=None, komi=komi)           -> ``Position.__init__``
  go_env.py:125-128  board == BLACK / WHITE planes                     -> ``board`` (+1 black, -1 white)

and the wrapper itself (``GoEnvOracle`` below follows go_env.py:38-88,156-230,339-361 line by
line: 16 history planes shifted by two per move in (mover, opponent) order, a constant player
plane, N*N+1 actions with pass = N*N, rewards [1,-1] if black wins else [-1,1]).

The reference's only tests at this boundary (``test_go_env.py:21-23``) assert types, so there is no
golden vector to pin against: the known-answer positions in ``tests/test_go_oracle.py`` (capture,
ko, suicide, seki/dame scoring) are hand-derived from the rules.  Deliberately simple: groups are
re-discovered by flood fill on every call instead of MiniGo's incremental liberty tracker.
"""
import copy

import numpy as np

BLACK, WHITE, EMPTY = 1, -1, 0


class IllegalMove(Exception):
    pass


def from_flat(n, flat):
    """coords.from_flat: action -> (row, col), or None for the pass action N*N."""
    if flat == n * n:
        return None
    return divmod(int(flat), n)


def _neighbors(n, c):
    r, w = c
    out = []
    for rr, ww in ((r + 1, w), (r - 1, w), (r, w + 1), (r, w - 1)):
        if 0 <= rr < n and 0 <= ww < n:
            out.append((rr, ww))
    return out


def find_reached(board, c):
    """The connected same-colour region containing c and the set of points bordering it."""
    n = board.shape[0]
    color = board[c]
    chain = {c}
    reached = set()
    frontier = [c]
    while frontier:
        cur = frontier.pop()
        for nb in _neighbors(n, cur):
            if board[nb] == color:
                if nb not in chain:
                    chain.add(nb)
                    frontier.append(nb)
            else:
                reached.add(nb)
    return chain, reached


def is_koish(board, c):
    """Colour that surrounds the empty point c on every side, else None."""
    if board[c] != EMPTY:
        return None
    colors = {int(board[nb]) for nb in _neighbors(board.shape[0], c)}
    if len(colors) == 1 and EMPTY not in colors:
        return colors.pop()
    return None


class Position(object):
    """go_base.Position restated: ``board`` int8 [N,N] (+1 black, -1 white), ``ko`` point or None,
    ``to_play``, ``recent`` = tuple of (colour, move) with move None for a pass, ``n`` moves played."""

    def __init__(self, n=19, komi=7.5, board=None):
        self.N = n
        self.board = np.zeros((n, n), dtype=np.int8) if board is None else np.array(board, dtype=np.int8)
        self.komi = komi
        self.ko = None
        self.to_play = BLACK
        self.recent = ()
        self.n = 0
        self.caps = (0, 0)

    def _liberties(self, chain_border):
        return {p for p in chain_border if self.board[p] == EMPTY}

    def is_move_suicidal(self, move):
        potential_libs = set()
        for nb in _neighbors(self.N, move):
            if self.board[nb] == EMPTY:
                return False                       # a liberty of its own
            chain, border = find_reached(self.board, nb)
            libs = self._liberties(border)
            if self.board[nb] == self.to_play:
                potential_libs |= libs
            elif len(libs) == 1:
                return False                       # captures that opponent group
        potential_libs -= {move}
        return not potential_libs

    def is_move_legal(self, move):
        if move is None:
            return True
        if self.board[move] != EMPTY:
            return False
        if move == self.ko:
            return False
        if self.is_move_suicidal(move):
            return False
        return True

    def all_legal_moves(self):
        """int8 [N*N + 1]: 1 = legal; the pass (last entry) is always legal."""
        n = self.N
        legal = np.zeros(n * n + 1, dtype=np.int8)
        for r in range(n):
            for w in range(n):
                if self.is_move_legal((r, w)):
                    legal[r * n + w] = 1
        legal[n * n] = 1
        return legal

    def pass_move(self):
        pos = copy.deepcopy(self)
        pos.n += 1
        pos.recent += ((pos.to_play, None),)
        pos.to_play = -pos.to_play
        pos.ko = None
        return pos

    def play_move(self, c):
        """Returns the NEW position (the reference rebinds ``self._go``, go_env.py:172)."""
        if c is None:
            return self.pass_move()
        if not self.is_move_legal(c):
            raise IllegalMove('%s move at %r is illegal' % ('Black' if self.to_play == BLACK else 'White', c))
        pos = copy.deepcopy(self)
        color = pos.to_play
        potential_ko = is_koish(pos.board, c)
        pos.board[c] = color
        captured = set()
        for nb in _neighbors(pos.N, c):
            if pos.board[nb] == -color:
                chain, border = find_reached(pos.board, nb)
                if not any(pos.board[p] == EMPTY for p in border):
                    captured |= chain
                    for p in chain:
                        pos.board[p] = EMPTY
        if len(captured) == 1 and potential_ko == -color:
            pos.ko = next(iter(captured))
        else:
            pos.ko = None
        if color == BLACK:
            pos.caps = (pos.caps[0] + len(captured), pos.caps[1])
        else:
            pos.caps = (pos.caps[0], pos.caps[1] + len(captured))
        pos.n += 1
        pos.recent += ((color, c),)
        pos.to_play = -color
        return pos

    def is_game_over(self):
        return len(self.recent) >= 2 and self.recent[-1][1] is None and self.recent[-2][1] is None

    def score(self):
        """Tromp-Taylor area score from black's point of view, komi subtracted."""
        work = np.array(self.board, dtype=np.int8)
        unknown = 2
        n = self.N
        for r in range(n):
            for w in range(n):
                if work[r, w] == EMPTY:
                    territory, border = find_reached(work, (r, w))
                    colors = {int(work[b]) for b in border}
                    if BLACK in colors and WHITE not in colors:
                        fill = BLACK
                    elif WHITE in colors and BLACK not in colors:
                        fill = WHITE
                    else:
                        fill = unknown               # dame or seki
                    for p in territory:
                        work[p] = fill
        return int(np.count_nonzero(work == BLACK)) - int(np.count_nonzero(work == WHITE)) - self.komi

    def result(self):
        s = self.score()
        return 1 if s > 0 else (-1 if s < 0 else 0)


class GoEnvOracle(object):
    """go_env.py:30-373 restated without PettingZoo / pygame (rendering omitted)."""

    def __init__(self, board_size=19, komi=7.5):
        self._N = board_size
        self._komi = komi
        self.agents = ['black_0', 'white_0']
        self.possible_agents = self.agents[:]
        self.board_history = np.zeros((board_size, board_size, 16), dtype=bool)
        self._is_terminal = False
        self.current_player_index = 0

    # go_env.py:212-230
    def reset(self, seed=None, options=None):
        self._go = Position(self._N, self._komi)
        self.agent_selection = 'black_0'
        self._cumulative_rewards = {'black_0': np.float64(0.0), 'white_0': np.float64(0.0)}
        self.rewards = {'black_0': np.float64(0.0), 'white_0': np.float64(0.0)}
        self.terminations = {'black_0': False, 'white_0': False}
        self.infos = {'black_0': {}, 'white_0': {}}
        self.next_legal_moves = np.where(self._go.all_legal_moves() == 1)[0]
        self.board_history = np.zeros((self._N, self._N, 16), dtype=bool)
        self.current_player_index = 0
        self._is_terminal = False

    # go_env.py:156-166
    def observe(self, agent):
        plane = np.zeros((self._N, self._N), dtype=bool) if agent == 'black_0' else \
            np.ones((self._N, self._N), dtype=bool)
        observation = np.dstack((self.board_history, plane))
        legal = self.next_legal_moves if agent == self.agent_selection else []
        mask = np.zeros(self._N * self._N + 1, 'int8')
        for i in legal:
            mask[i] = 1
        return {'observation': observation, 'action_mask': mask}

    # go_env.py:168-210
    def step(self, action):
        mover = self.agent_selection
        if self.terminations[mover]:
            raise RuntimeError('step() after the game ended (PettingZoo _was_dead_step)')
        self._go = self._go.play_move(from_flat(self._N, action))
        factor = BLACK if mover == 'black_0' else WHITE
        cur = self._go.board == factor
        opp = self._go.board == -factor
        self.board_history = np.dstack((cur, opp, self.board_history[:, :, :-2]))
        nxt = 'white_0' if mover == 'black_0' else 'black_0'
        self.current_player_index = self.agents.index(nxt)
        if self._go.is_game_over():
            self._is_terminal = True
            self.terminations = {'black_0': True, 'white_0': True}
            rw = [1, -1] if self._go.result() == 1 else [-1, 1]       # go_env.py:142-143
            self.rewards = {'black_0': rw[0], 'white_0': rw[1]}
            self.next_legal_moves = [self._N * self._N]
        else:
            self.next_legal_moves = np.where(self._go.all_legal_moves() == 1)[0]
        self.agent_selection = nxt
        for a in self.agents:                                          # AECEnv._accumulate_rewards
            self._cumulative_rewards[a] = self._cumulative_rewards[a] + self.rewards[a]
        obs = self.observe(nxt)
        return obs, self._cumulative_rewards[nxt], self.terminations[nxt], self.infos[nxt]

    def current_player(self):
        return self.current_player_index

    to_play = current_player

    def legal_actions(self, agent=None):
        return self.next_legal_moves

    def max_utility(self):
        return 1

    def returns(self):
        return [self.rewards['black_0'], self.rewards['white_0']]

    def clone(self):
        return copy.deepcopy(self)

    def is_terminal(self):
        return self._is_terminal


class GoSearchBoard(object):
    """The env duck-type ``pyoracle.Search`` (= the reference's ``AlphaZeroMCTS``,
    alphazero_mcts.py:42-71) drives -- ``step / game_end_winner / current_player / leagel_actions``,
    plus the ``states`` / ``last_move`` attributes the closed-form evaluators hash -- over a Go
    ``Position``.  The reference itself never runs AlphaZeroMCTS on GoEnv (its Go path is
    DeepMindMCTS); this adapter is what BASELINE.json's config 4 (AlphaZero on a Go board) means in
    the reference's own search.  Actions ascend with the pass (N*N) last; players 0 = black, 1 = white;
    ``max_moves`` > 0 ends and scores the game after that many moves (the engine's cap)."""

    def __init__(self, board_size=9, komi=7.5, max_moves=0):
        self.board_size = board_size
        self.komi = komi
        self.max_moves = max_moves
        self.reset()

    def reset(self):
        self.pos = Position(self.board_size, self.komi)
        self.last_move = -1

    def step(self, action):
        self.pos = self.pos.play_move(from_flat(self.board_size, action))
        self.last_move = int(action)

    @property
    def states(self):
        n = self.board_size
        return {int(r * n + c): (0 if self.pos.board[r, c] == BLACK else 1)
                for r, c in zip(*np.nonzero(self.pos.board))}

    def leagel_actions(self):
        return [int(a) for a in np.where(self.pos.all_legal_moves() == 1)[0]]

    def current_player(self):
        return 0 if self.pos.to_play == BLACK else 1

    def game_end_winner(self):
        if self.pos.is_game_over() or (self.max_moves and self.pos.n >= self.max_moves):
            return True, (0 if self.pos.result() == 1 else 1)      # go_env.py:142-143
        return False, -1
