"""``AlphaZeroMCTS`` / ``AlphaZeroPlayer`` with the reference's API
(rlzero/mcts/alphazero_mcts.py:17-165), searching on the GPU.

The object owns a one-tree ``SearchForest``; ``simulate`` uploads the caller's position and
runs ``n_playout`` waves of select -> evaluate -> expand+backup kernels.  Evaluation goes

* through the device when ``policy_value_fn`` is a native evaluator (it carries a
  ``device_evaluator`` attribute, e.g. ``AlphaZeroAgent.policy_value_fn``), or
* through the Python callable itself (``policy_value_fn(game_env)`` per leaf, the
  reference's contract, :27-31) -- slow, but it accepts any user evaluator.

The final ``softmax(log(N + 1e-10) / T)`` and the move sampling are done on the host with
numpy exactly as the reference does (:88-92, :148), so for the same visit counts and the same
global ``np.random`` state they return the same ``act_probs`` and the same move.
"""
import numpy as np
import torch

from .. import _lib as L
from ..engine import HostCallbackEvaluator, SearchForest
from ..games.gomoku.gomoku_env import LeafEnvView
from .node import TreeNode
from .player import Player


def softmax(x):
    """avoid data overflow (alphazero_mcts.py:10-14)."""
    probs = np.exp(x - np.max(x))
    probs /= np.sum(probs)
    return probs


class _NumpyNoiseCallback(HostCallbackEvaluator):
    """Host-callback evaluator that also draws the per-node Dirichlet noise from the global
    numpy RNG in the reference's order (node.py:63-69), keeping the RNG stream -- and hence
    every later ``np.random.choice`` -- aligned with the reference.  The mix is formed with the
    reference's own expression on the evaluator's own number types, so a forest that keeps
    float64 priors (PUCT rule) holds ``TreeNode.prior`` bit for bit."""

    def __init__(self, policy_value_fn, env_factory, add_noise):
        super().__init__(policy_value_fn, env_factory)
        self.add_noise = add_noise

    def _mix(self, act_probs, meta):
        act_probs = list(act_probs)
        if not self.add_noise:
            return act_probs
        noise = np.random.dirichlet(0.3 * np.ones(len(act_probs)))
        return [(a, 0.75 * p + 0.25 * noise[i]) for i, (a, p) in enumerate(act_probs)]


class AlphaZeroMCTS(object):
    """An implementation of Monte Carlo Tree Search (GPU trees, reference API)."""

    def __init__(self, policy_value_fn, n_playout=1000, c_puct=5, add_noise=False,
                 rule=L.RULE_UCT, device='cuda', leaves_per_wave=1, virtual_loss=1.0):
        """``leaves_per_wave = K > 1`` (an extension, not in the reference): leaf-parallel search with virtual loss --
        K playouts of this one tree share a network batch.  Faster for a single game, but no longer the reference's
        sequential playout order (visit counts differ); the default 1 is bit-exact."""
        self.leaves_per_wave = max(1, int(leaves_per_wave))
        self.virtual_loss = float(virtual_loss)
        self.policy_value_fn = policy_value_fn
        self.n_playout = n_playout
        self._c_puct = c_puct
        self.add_noise = add_noise
        self.rule = rule
        self.device = device
        self._forest = None
        self._evaluator = None
        self._synced = False      # forest root position == last env passed to simulate
        self._snap = None
        self._seed = 0

    # ---------------------------------------------------------------- plumbing
    def _ensure_forest(self, env):
        f = self._forest
        width = getattr(env, 'board_width', env.board_size)
        game_type = getattr(env, 'game_type', L.GAME_GOMOKU)
        if (f is not None and f.H == env.board_size and f.W == width and f.game_type == game_type
                and f.k == env.n_in_row and self.n_playout <= f.n_playout and f.c_puct == float(self._c_puct)
                and f.K == self.leaves_per_wave):
            return f     # n_playout may be lowered between moves (the pools were sized for the larger one)
        native = getattr(self, '_native_evaluator', None)
        if native is None:
            native = getattr(self.policy_value_fn, 'device_evaluator', None)
        # update_with_move keeps the whole subtree of the move played (alphazero_mcts.py:96-103), which a search
        # can grow to any size over a game: one tree costs little, so the pool takes everything the re-root kernel
        # can compact (6144 nodes); a kept subtree beyond that restarts the tree (counted in forest.carry_dropped)
        carry = max(0, min(6144 - int(self.n_playout), max(4 * int(self.n_playout), 1024)))
        self._forest = SearchForest(1, env.board_size, env.n_in_row, n_playout=self.n_playout,
                                    c_puct=self._c_puct, rule=self.rule, max_carry=carry,
                                    device=self.device, board_width=width, game_type=game_type,
                                    leaves_per_tree=self.leaves_per_wave, virtual_loss=self.virtual_loss,
                                    prior_f64=(self.rule == L.RULE_PUCT and native is None))
        if native is not None:
            self._evaluator = native
        else:
            h, k = env.board_size, env.n_in_row
            self._evaluator = _NumpyNoiseCallback(
                self.policy_value_fn, lambda rows, meta: LeafEnvView(rows, meta, h, k, width, game_type),
                self.add_noise)
        return self._forest

    def _upload(self, env):
        f = self._forest
        rows, meta = env.device_state()
        f.root_rows.copy_(rows.to(f.device))
        keep = f.root_meta.clone()
        f.root_meta.copy_(meta.to(f.device))
        f.root_meta[:, L.META_FAULT] = 0
        f.root_meta[:, L.META_EPISODE] = keep[:, L.META_EPISODE]
        f.root_meta[:, L.META_STATUS] = L.ACTIVE  # the reference searches whatever it is given
        self._synced = True

    @property
    def _root(self):
        """Host snapshot of the tree as reference-style ``TreeNode`` objects."""
        if self._forest is None:
            return TreeNode(None, 1.0)
        if self._snap is None:
            self._snap = TreeNode.from_snapshot(self._forest.dump_tree(0))
        return self._snap

    # ------------------------------------------------------------ reference API
    def _playout(self, game_env):
        """One playout from the root (alphazero_mcts.py:42-71).  ``game_env`` is only read."""
        f = self._ensure_forest(game_env)
        self._upload(game_env)
        self._waves(1)

    def _waves(self, n):
        f = self._forest
        native = getattr(self._evaluator, 'graph_capturable', False)
        eps = 0.25 if (self.add_noise and native) else 0.0  # callback path mixes noise on the host
        self._seed += 1
        f.search(self._evaluator, n, noise_eps=eps, noise_alpha=0.3, seed=self._seed)
        self._snap = None
        f.raise_faults()

    def simulate(self, game_env, temperature=1e-3):
        """Run ``n_playout`` playouts from ``game_env``'s position; return the root's actions
        and ``softmax(1/T * log(visits + 1e-10))`` (alphazero_mcts.py:73-94)."""
        self._ensure_forest(game_env)
        self._upload(game_env)
        self._waves(self.n_playout)
        visits, _, has, _, _ = self._forest.root_stats()
        acts = tuple(int(a) for a in np.nonzero(has[0])[0])
        if not acts:
            raise ValueError('not enough values to unpack (expected 2, got 0)')  # :90 on a terminal root
        act_probs = softmax(1.0 / temperature * np.log(visits[0][list(acts)].astype(np.int64) + 1e-10))
        return acts, act_probs

    def update_with_move(self, last_move):
        """Step forward in the tree, keeping the subtree below ``last_move`` if the root has
        that child, else start a fresh root (alphazero_mcts.py:96-103)."""
        f = self._forest
        self._snap = None
        if f is None:
            return
        keep = False
        if last_move is not None and last_move >= 0 and self._synced and int(f.n_nodes[0]) > 0:
            n = int(f.edge_N[int(last_move)]) if last_move < f.A else -1
            keep = n >= 0
        if keep:
            f.advance([int(last_move)], keep_subtree=True)
            f.raise_faults()
        else:
            f.advance([-1])
        self._synced = False

    def __str__(self):
        return 'AlphaZeroMCTS'


class AlphaZeroPlayer(Player):
    """AI player based on MCTS (alphazero_mcts.py:109-165)."""

    def __init__(self, policy_value_fn, n_playout=1000, c_puct=5, is_selfplay=False,
                 player_id=0, player_name='', rule=L.RULE_UCT, device='cuda', leaves_per_wave=1,
                 virtual_loss=1.0):
        super().__init__(player_id, player_name)
        self.is_selfplay = is_selfplay
        self.add_noise = is_selfplay
        self.mcts = AlphaZeroMCTS(policy_value_fn, n_playout=n_playout, c_puct=c_puct,
                                  add_noise=self.add_noise, rule=rule, device=device,
                                  leaves_per_wave=leaves_per_wave, virtual_loss=virtual_loss)

    def reset_player(self):
        """reset, reconstructing the MCTS Tree for next simulation."""
        self.mcts.update_with_move(-1)

    def get_action(self, game_env, temperature=1e-3, return_prob=False):
        sensible_moves = game_env.leagel_actions()
        # the pi vector returned by MCTS as in the alphaGo Zero paper
        # board_size**2 in the reference (:144); envs with another action set say so via n_actions
        move_probs = np.zeros(getattr(game_env, 'n_actions', game_env.board_size * game_env.board_size))
        if len(sensible_moves) == 0:
            print('WARNING: the board is full')
            return None
        acts, probs = self.mcts.simulate(game_env, temperature)
        move_probs[list(acts)] = probs
        move = np.random.choice(acts, p=probs)
        if self.is_selfplay:
            self.mcts.update_with_move(move)       # keep the subtree (:149-152)
        else:
            move = np.random.choice(acts, p=probs)  # the reference samples a second time (:157)
            self.mcts.update_with_move(-1)
        return (move, move_probs) if return_prob else move

    def __str__(self):
        return 'AlphaZeroPlayer, id: {}, name: {}.'.format(self.get_player_id(), self.get_player_name())
