"""``GomokuEnv`` with the reference's API (rlzero/games/gomoku/gomoku_env.py:11-285), rules on
the GPU.

The position lives in a 1-game device record (row bitmasks + meta, the layout of
``include/rlzero_b200.h``) and every rule -- placing a stone, the k-in-a-row test, terminal
detection, the 4-plane observation -- is computed by the same CUDA kernels the batched
engine uses (``rz_gomoku_step / _winner / _encode_f32``).  The host object only mirrors the
bookkeeping the reference exposes as attributes (``states`` dict in play order, the ascending
``leagel_actions`` list, ``last_move``).  No CUDA library, no env: there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from ... import _lib as L
from ..base_env import BaseEnv


class Error(Exception):
    """Stand-in for ``uu.Error`` raised by the reference on bad reset arguments
    (gomoku_env.py:4,35-40; the ``uu`` module is gone in Python 3.13)."""


class GomokuEnv(BaseEnv):
    """board for the game (k-in-a-row on an N x N board)."""

    def __init__(self, board_size=8, n_in_row=5, start_player_idx=0, device='cuda'):
        super().__init__()
        self.board_size = board_size
        self.n_in_row = n_in_row
        self.players = [0, 1]
        self.start_player_idx = start_player_idx
        self._current_player = self.players[self.start_player_idx]
        # geometry extension (SURVEY 7 "square-board assumption"): square k-in-a-row by default
        self.board_width = board_size
        self.game_type = L.GAME_GOMOKU
        self.n_actions = self.board_size * self.board_width
        self._leagel_actions = list(range(self.n_actions))
        self._device = device
        self._rows = None
        self._meta = None

    # ------------------------------------------------------------ device side
    def _gdesc(self):
        a = self.n_actions
        return L.GameDesc(self.board_size, self.n_in_row, a, (a + 31) // 32 * 32, self.board_width,
                          self.game_type)

    def _ensure_device(self):
        if self._rows is None:
            if not torch.cuda.is_available():
                raise L.NativeLibraryError('GomokuEnv needs a CUDA device (no CPU fallback)')
            if not (1 <= self.board_size <= L.MAX_BOARD and 1 <= self.board_width <= L.MAX_BOARD):
                raise Error('board_size must be in [1, %d]' % L.MAX_BOARD)
            self._lib = L.load()
            self._rows = torch.zeros(1, 2, self.board_size, dtype=torch.int32, device=self._device)
            self._meta = torch.zeros(1, L.META_STRIDE, dtype=torch.int32, device=self._device)

    def device_state(self):
        """(rows[1,2,H] int32, meta[1,8] int32) CUDA tensors of the current position."""
        return self._rows, self._meta

    def __deepcopy__(self, memo):
        new = type(self).__new__(type(self))
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                setattr(new, k, v.clone())
            elif isinstance(v, (list, dict)):
                setattr(new, k, type(v)(v))
            else:
                setattr(new, k, v)
        return new

    # ------------------------------------------------------------- reference API
    def reset(self, start_player_idx=0):
        """init the board and set some variables (gomoku_env.py:33-47)."""
        if max(self.board_size, self.board_width) < self.n_in_row:
            raise Error(f'Board board_size can not less than {self.n_in_row}')
        if start_player_idx not in (0, 1):
            raise Error(f'{start_player_idx} should be 0 (player1 first) or 1 (player2 first)')
        self._ensure_device()
        self.start_player_idx = start_player_idx
        self._current_player = self.players[start_player_idx]
        self._leagel_actions = list(range(self.n_actions))
        self.states = {}
        self.last_move = -1
        self.info = {}
        g = self._gdesc()
        L.check(self._lib.rz_gomoku_reset(C.byref(g), L.ptr(self._rows), L.ptr(self._meta), 1, 0,
                                          L.stream_ptr()), 'rz_gomoku_reset')
        if start_player_idx:
            self._meta[0, L.META_PLAYER] = start_player_idx
        return self.current_state()

    def step(self, action):
        """Update the board (gomoku_env.py:49-70)."""
        assert (action in self._leagel_actions), print(
            f'You input illegal action: {action}, the legal_actions are {self._leagel_actions}.')
        g = self._gdesc()
        act = torch.tensor([int(action)], dtype=torch.int32, device=self._rows.device)
        out = torch.zeros(2, dtype=torch.int32, device=self._rows.device)
        L.check(self._lib.rz_gomoku_step(C.byref(g), L.ptr(self._rows), L.ptr(self._meta), L.ptr(act),
                                         L.ptr(out[0:1]), L.ptr(out[1:2]), 1, L.stream_ptr()),
                'rz_gomoku_step')
        reward, win = (int(x) for x in out.cpu().numpy())
        self._record_move(action)
        self._current_player = (self.players[0] if self._current_player == self.players[1]
                                else self.players[1])
        return self.current_state(), reward, bool(win), self.info

    def _record_move(self, action):
        """host mirror of the move: states dict, legal list, last_move (gomoku_env.py:55-57)."""
        self.states[action] = self._current_player
        self._leagel_actions.remove(action)
        self.last_move = action

    def leagel_actions(self):
        return self._leagel_actions

    def legal_actions(self, player=None):
        return self._leagel_actions

    def current_state(self):
        """4 x N x N float64 planes from the mover's perspective (gomoku_env.py:95-114)."""
        n = self.board_size
        out = torch.empty(1, 4, n, self.board_width, dtype=torch.float32, device=self._rows.device)
        g = self._gdesc()
        L.check(self._lib.rz_gomoku_encode_f32(C.byref(g), L.ptr(self._rows), L.ptr(self._meta),
                                               L.ptr(out), 1, L.stream_ptr()), 'rz_gomoku_encode_f32')
        return out[0].cpu().numpy().astype(np.float64)

    def _end_winner(self):
        g = self._gdesc()
        out = torch.zeros(2, dtype=torch.int32, device=self._rows.device)
        L.check(self._lib.rz_gomoku_winner(C.byref(g), L.ptr(self._rows), L.ptr(self._meta),
                                           L.ptr(out[0:1]), L.ptr(out[1:2]), 1, L.stream_ptr()),
                'rz_gomoku_winner')
        end, winner = (int(x) for x in out.cpu().numpy())
        return bool(end), winner

    def has_a_winner(self):
        """(True, player) if someone has n_in_row in a line, else (False, -1) (gomoku_env.py:116-170)."""
        _, winner = self._end_winner()
        return (winner != -1), winner

    def game_end_winner(self):
        """(gomoku_env.py:196-203)"""
        end, winner = self._end_winner()
        return (True, winner) if end else (False, -1)

    def is_terminal(self):
        return self.game_end_winner()[0]

    def get_done_reward(self):
        """(gomoku_env.py:172-194) -- keeps the reference's ``winner == 1 / == 2`` tests although
        players are numbered 0/1."""
        win, winner = self.has_a_winner()
        reward = None
        if winner == 1:
            reward = 1
        elif winner == 2:
            reward = -1
        elif winner == -1 and win:
            reward = 0
        return win, reward

    def returns(self):
        """(gomoku_env.py:210-225) -- same 1/2 quirk: a player-0 win reports [0, 0]."""
        _, winner = self.has_a_winner()
        if winner == 1:
            return [1, -1]
        if winner == 2:
            return [-1, 1]
        return [0, 0]

    def move_to_location(self, move):
        return [move // self.board_size, move % self.board_size]

    def location_to_move(self, location):
        if len(location) != 2:
            return -1
        move = location[0] * self.board_size + location[1]
        if move not in range(self.board_size * self.board_size):
            return -1
        return move

    def action_to_string(self, move):
        return f'Play row {move // self.board_size + 1}, column {move % self.board_size + 1}'

    def max_utility(self):
        return 1

    def current_player(self):
        return self._current_player

    def current_player_index(self):
        return 0 if self._current_player == 1 else 1

    def render(self):
        n = self.board_size
        print()
        for x in range(n):
            print('{0:8}'.format(x), end='')
        print('\r\n')
        for i in range(n - 1, -1, -1):
            print('{0:4d}'.format(i), end='')
            for j in range(n):
                p = self.states.get(i * n + j, -1)
                print(('B' if p == 0 else 'W' if p == 1 else '_').center(8), end='')
            print('\r\n\r\n')

    def __str__(self):
        return 'Gomoku Board'


class LeafEnvView(object):
    """Host view of one leaf position produced by ``rz_tree_select`` -- what a user-supplied
    ``policy_value_fn(game_env)`` receives on the slow path (alphazero_mcts.py:59).  It offers
    the env attributes evaluators use: ``current_state()``, ``leagel_actions()``, ``states``,
    ``last_move``, ``current_player()``, ``board_size``, ``game_end_winner()``.  ``states``
    lists stones in ascending square order (play order below the root is not recorded)."""

    def __init__(self, rows, meta, board_size, n_in_row, board_width=None, game_type=L.GAME_GOMOKU):
        self.board_size = board_size
        self.board_width = board_size if board_width is None else board_width
        self.game_type = game_type
        self.n_in_row = n_in_row
        self.players = [0, 1]
        self._rows = np.asarray(rows, dtype=np.uint32)
        self._meta = np.asarray(meta)
        self.last_move = int(meta[L.META_LAST_MOVE])
        self._current_player = int(meta[L.META_PLAYER])
        n, w = board_size, self.board_width
        bits = ((self._rows[:, :, None] >> np.arange(w, dtype=np.uint32)[None, None, :]) & 1).astype(bool)
        self._bits = bits  # [2, n, w]
        occ = bits[0] | bits[1]
        if game_type == L.GAME_CONNECT4:
            self.n_actions = w
            self._leagel_actions = np.nonzero(~occ[n - 1])[0].tolist()      # columns whose top square is empty
        else:
            self.n_actions = n * w
            self._leagel_actions = np.nonzero(~occ.reshape(-1))[0].tolist()
        self.states = {}
        who = np.where(bits[0], 0, 1).reshape(-1)
        for m in np.nonzero(occ.reshape(-1))[0].tolist():
            self.states[m] = int(who[m])

    def leagel_actions(self):
        return self._leagel_actions

    def legal_actions(self, player=None):
        return self._leagel_actions

    def current_player(self):
        return self._current_player

    def current_state(self):
        n, w = self.board_size, self.board_width
        planes = np.zeros((4, n, w))
        planes[0] = self._bits[self._current_player]
        planes[1] = self._bits[1 - self._current_player]
        if self.states:
            planes[2, self.last_move // w, self.last_move % w] = 1.0
        if len(self.states) % 2 == 0:
            planes[3] = 1.0
        return planes

    def game_end_winner(self):
        st = int(self._meta[L.META_STATUS])
        if st == L.ENDED_WIN:
            return True, int(self._meta[L.META_WINNER])
        if st == L.ENDED_TIE:
            return True, -1
        return False, -1
