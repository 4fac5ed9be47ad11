"""``ConnectFourEnv``: 6x7 Connect Four behind the duck-typed env API the reference's search and
game loop use (rlzero/games/base_env.py:7-33, the GomokuEnv method set of
rlzero/games/gomoku/gomoku_env.py:11-285).

The reference has no Connect Four env; BASELINE.json config 2 asks for one.
It is defined here as k-in-a-row WITH GRAVITY so that everything else -- bitboard layout, win
test, observation planes, kernels -- is shared with Gomoku (``RZ_GAME_CONNECT4`` in
include/rlzero_b200.h):

* an action is a COLUMN 0..W-1; the stone drops to the lowest empty row (row 0 = bottom);
* ``leagel_actions()`` = ascending list of columns that are not full (a full column leaves the
  list, like a played square leaves GomokuEnv's list, gomoku_env.py:56);
* ``states`` maps SQUARE r*W+c -> player, ``last_move`` is the square of the last stone, so
  ``current_state()`` is the same 4 planes (mover's stones, opponent's, last move, colour to move);
* rewards / terminal rule / player ids as GomokuEnv (gomoku_env.py:49-70,196-203).
"""
from ... import _lib as L
from ..gomoku.gomoku_env import GomokuEnv


class ConnectFourEnv(GomokuEnv):

    def __init__(self, rows=6, cols=7, n_in_row=4, start_player_idx=0, device='cuda'):
        super().__init__(board_size=rows, n_in_row=n_in_row, start_player_idx=start_player_idx, device=device)
        self.board_width = cols
        self.game_type = L.GAME_CONNECT4
        self.n_actions = cols
        self._leagel_actions = list(range(cols))
        self._heights = [0] * cols

    def reset(self, start_player_idx=0):
        self._heights = [0] * self.board_width
        return super().reset(start_player_idx)

    def _record_move(self, action):
        cell = self._heights[action] * self.board_width + action
        self._heights[action] += 1
        self.states[cell] = self._current_player
        if self._heights[action] >= self.board_size:
            self._leagel_actions.remove(action)
        self.last_move = cell

    def move_to_location(self, move):
        return [self._heights[move], move]

    def action_to_string(self, move):
        return f'Drop in column {move + 1}'

    def __str__(self):
        return 'Connect Four Board'
