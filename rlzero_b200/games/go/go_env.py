"""``GoEnv`` with the reference's API (rlzero/games/go/go_env.py:30-373), rules on the GPU.

The reference adapts PettingZoo's ``go_v5`` (``pettingzoo.classic.go.go_base`` = MiniGo's rules) to
its ``BaseEnv`` interface.  Here the position lives in device records (row bitmasks of both colours,
14 history planes, meta incl. ko point and pass count; ``include/rlzero_b200.h``) and every rule --
captures, suicide and ko legality, the two-pass end, Tromp-Taylor scoring, the 17-plane observation
-- is computed by the CUDA kernels of ``csrc/rz_go.cu`` (``rz_go_step / _legal_mask / _score /
_encode_f32``), the same device functions the batched search descends with.

``GoBoards`` is the batched form (n games per call); ``GoEnv`` is a 1-game view of it exposing the
reference's methods and attributes.  PettingZoo plumbing that has no meaning off-screen (render,
pygame surfaces, ``observation_spaces`` gym objects) is not reproduced.  No CUDA library, no env.
"""
import copy
import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from ... import _lib as L

BaseEnvTimestep = namedtuple('BaseEnvTimestep', ['obs', 'reward', 'done', 'info'])   # go_env.py:18-19


class GoBoards(object):
    """n Go positions resident on the device + the batched rule kernels."""

    def __init__(self, n, board_size=19, komi=7.5, max_moves=0, device='cuda'):
        if not torch.cuda.is_available():
            raise L.NativeLibraryError('GoBoards needs a CUDA device (no CPU fallback)')
        if not 1 <= board_size <= L.MAX_BOARD:
            raise ValueError('board_size must be in [1, %d]' % L.MAX_BOARD)
        self.lib = L.load()
        self.n, self.N = int(n), int(board_size)
        self.cells = self.N * self.N
        self.A = self.cells + 1
        self.komi = float(komi)
        self.device = torch.device(device)
        self.gdesc = L.GameDesc(self.N, 1, self.A, (self.A + 31) // 32 * 32, self.N, L.GAME_GO, self.komi,
                                int(max_moves))
        i32 = torch.int32
        self.rows = torch.zeros(self.n, 2, self.N, dtype=i32, device=self.device)
        self.hist = torch.zeros(self.n, L.GO_HIST, self.N, dtype=i32, device=self.device)
        self.meta = torch.zeros(self.n, L.META_STRIDE, dtype=i32, device=self.device)
        self.reward = torch.zeros(self.n, 2, dtype=i32, device=self.device)
        self.done = torch.zeros(self.n, dtype=i32, device=self.device)
        self.reset()

    def _s(self):
        return L.stream_ptr()

    def reset(self, only_ended=False):
        L.check(self.lib.rz_go_reset(C.byref(self.gdesc), L.ptr(self.rows), L.ptr(self.hist), L.ptr(self.meta),
                                     self.n, int(only_ended), self._s()), 'rz_go_reset')

    def step(self, actions):
        """actions int [n] (-1 skips a game) -> (reward int32 [n,2], done int32 [n]) device tensors."""
        a = torch.as_tensor(actions, dtype=torch.int32, device=self.device).contiguous()
        L.check(self.lib.rz_go_step(C.byref(self.gdesc), L.ptr(self.rows), L.ptr(self.hist), L.ptr(self.meta),
                                    L.ptr(a), L.ptr(self.reward), L.ptr(self.done), self.n, self._s()),
                'rz_go_step')
        return self.reward, self.done

    def legal_mask(self):
        m = torch.zeros(self.n, self.A, dtype=torch.uint8, device=self.device)
        L.check(self.lib.rz_go_legal_mask(C.byref(self.gdesc), L.ptr(self.rows), L.ptr(self.meta), L.ptr(m),
                                          self.n, self._s()), 'rz_go_legal_mask')
        return m

    def score(self):
        s = torch.zeros(self.n, dtype=torch.float64, device=self.device)
        r = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        L.check(self.lib.rz_go_score(C.byref(self.gdesc), L.ptr(self.rows), L.ptr(self.meta), L.ptr(s), L.ptr(r),
                                     self.n, self._s()), 'rz_go_score')
        return s, r

    def observe(self):
        """float32 [n,17,N,N]: 16 history planes + player plane (go_env.py:156-166, channels first)."""
        p = torch.zeros(self.n, 17, self.N, self.N, dtype=torch.float32, device=self.device)
        L.check(self.lib.rz_go_encode_f32(C.byref(self.gdesc), L.ptr(self.rows), L.ptr(self.hist),
                                          L.ptr(self.meta), L.ptr(p), self.n, self._s()), 'rz_go_encode_f32')
        return p

    def boards(self):
        """int8 [n,N,N] host array in go_base's convention: +1 black, -1 white, 0 empty."""
        rows = self.rows.cpu().numpy().view(np.uint32)
        bits = (rows[:, :, :, None] >> np.arange(self.N, dtype=np.uint32)[None, None, None, :]) & 1
        return (bits[:, 0].astype(np.int8) - bits[:, 1].astype(np.int8))

    def faults(self):
        return self.meta[:, L.META_FAULT].cpu().numpy()


class GoEnv(object):
    """go_env.py:30-373 behind the same method names; one game on the device."""
    metadata = {'render_modes': ['human', 'rgb_array'], 'name': 'go_v5', 'is_parallelizable': False,
                'render_fps': 2}

    def __init__(self, board_size=19, komi=7.5, render_mode=None, screen_height=800, device='cuda'):
        self._N = board_size
        self._komi = komi
        self.agents = ['black_0', 'white_0']
        self.possible_agents = self.agents[:]
        self.render_mode = render_mode
        self._device = device
        self._boards = None
        self._is_terminal = False
        self.current_player_index = 0
        self.board_history = np.zeros((self._N, self._N, 16), dtype=bool)

    # -------------------------------------------------------------- device side
    def _ensure_device(self):
        if self._boards is None:
            self._boards = GoBoards(1, self._N, self._komi, device=self._device)

    def device_state(self):
        b = self._boards
        return b.rows, b.hist, b.meta

    def __deepcopy__(self, memo):
        new = type(self).__new__(type(self))
        for k, v in self.__dict__.items():
            if k == '_boards' and v is not None:
                nb = copy.copy(v)
                for name in ('rows', 'hist', 'meta', 'reward', 'done'):
                    setattr(nb, name, getattr(v, name).clone())
                setattr(new, k, nb)
            else:
                setattr(new, k, copy.deepcopy(v, memo))
        return new

    def _refresh(self):
        """host mirrors of the device position: history planes and the legal-move list."""
        b = self._boards
        obs = b.observe()[0].cpu().numpy()
        self.board_history = np.ascontiguousarray(obs[:16].transpose(1, 2, 0)).astype(bool)
        if self._is_terminal:
            self.next_legal_moves = [self._N * self._N]                       # go_env.py:192
        else:
            self.next_legal_moves = np.where(b.legal_mask()[0].cpu().numpy() == 1)[0]

    # ------------------------------------------------------------ reference API
    def _convert_to_dict(self, list_of_list):
        return dict(zip(self.possible_agents, list_of_list))

    def reset(self, seed=None, options=None):
        """go_env.py:212-230."""
        self._ensure_device()
        self._boards.reset()
        self.agents = self.possible_agents[:]
        self.agent_selection = self.agents[0]
        self._cumulative_rewards = self._convert_to_dict(np.array([0.0, 0.0]))
        self.rewards = self._convert_to_dict(np.array([0.0, 0.0]))
        self.terminations = self._convert_to_dict([False, False])
        self.truncations = self._convert_to_dict([False, False])
        self.infos = self._convert_to_dict([{}, {}])
        self._is_terminal = False
        self.current_player_index = 0
        self._refresh()

    def observe(self, agent):
        """go_env.py:156-166."""
        plane = np.zeros([self._N, self._N], dtype=bool) if agent == self.possible_agents[0] else \
            np.ones([self._N, self._N], dtype=bool)
        observation = np.dstack((self.board_history, plane))
        legal_moves = self.next_legal_moves if agent == self.agent_selection else []
        action_mask = np.zeros((self._N * self._N) + 1, 'int8')
        for i in legal_moves:
            action_mask[i] = 1
        return {'observation': observation, 'action_mask': action_mask}

    def step(self, action):
        """go_env.py:168-210."""
        if self.terminations[self.agent_selection] or self.truncations[self.agent_selection]:
            raise RuntimeError('step() on a finished game (PettingZoo _was_dead_step)')
        b = self._boards
        reward, done = b.step([int(action)])
        if int(b.faults()[0]) & L.FAULT_ILLEGAL_MOVE:
            b.meta[:, L.META_FAULT] = 0
            raise ValueError('illegal move %r (go_base.IllegalMove)' % (action,))
        mover = self.agent_selection
        nxt = self.agents[1 - self.agents.index(mover)]
        self.current_player_index = self.agents.index(nxt)
        if int(done[0]):
            self._is_terminal = True
            self.terminations = self._convert_to_dict([True, True])
            self.rewards = self._convert_to_dict([int(x) for x in reward[0].cpu().numpy()])
        self.agent_selection = nxt
        for a in self.agents:                                              # AECEnv._accumulate_rewards
            self._cumulative_rewards[a] = self._cumulative_rewards[a] + self.rewards[a]
        self._refresh()
        agent = self.agent_selection
        return BaseEnvTimestep(self.observe(agent), self._cumulative_rewards[agent], self.terminations[agent],
                               self.infos[agent])

    def current_player(self):
        return self.current_player_index

    def to_play(self):
        return self.current_player_index

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

    def random_action(self):
        return np.random.choice(self.legal_actions())

    def seed(self, seed, dynamic_seed=True):
        self._seed = seed
        self._dynamic_seed = dynamic_seed
        np.random.seed(self._seed)

    def set_game_result(self, result_val):
        for i, name in enumerate(self.agents):
            self.terminations[name] = True
            self.rewards[name] = result_val * (1 if i == 0 else -1)
            self.infos[name] = {'legal_moves': []}

    def close(self):
        pass

    def __repr__(self):
        return 'LightZero Go Env'
