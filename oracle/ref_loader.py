"""TEST INFRASTRUCTURE ONLY -- import the unmodified reference from /root/reference.

The reference imports ``gymnasium`` only to subclass ``gymnasium.Env``
(``rlzero/games/base_env.py:4,7``); the module is absent in this image, so a
two-line stand-in is installed before the import.  Nothing else is patched.

``/root/reference`` exists only in the authoring container: ``available()``
is False on the GPU box and callers must then skip (tests) or use the
restatement in ``oracle.pyoracle`` (bench ``cpu_baseline``).
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get('RLZERO_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'rlzero', 'mcts'))


def load():
    """Return a namespace with the reference classes of the hot path."""
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    if 'gymnasium' not in sys.modules:
        g = types.ModuleType('gymnasium')
        g.Env = type('Env', (object,), {})
        sys.modules['gymnasium'] = g
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')  # 'uu' deprecation on py3.12
        from rlzero.games.gomoku import GameControl, GomokuEnv
        from rlzero.games.gomoku.alphazero_agent import AlphaZeroAgent
        from rlzero.games.gomoku.policy_value_net import PolicyValueNet
        from rlzero.mcts.alphazero_mcts import AlphaZeroMCTS, AlphaZeroPlayer
        from rlzero.mcts.node import TreeNode
        from rlzero.mcts.rollout_mcts import RolloutMCTS, RolloutPlayer
    return types.SimpleNamespace(
        GomokuEnv=GomokuEnv, GameControl=GameControl,
        AlphaZeroAgent=AlphaZeroAgent, PolicyValueNet=PolicyValueNet,
        AlphaZeroMCTS=AlphaZeroMCTS, AlphaZeroPlayer=AlphaZeroPlayer,
        TreeNode=TreeNode, RolloutMCTS=RolloutMCTS, RolloutPlayer=RolloutPlayer)


def use_puct_rule(ref):
    """Test-side switch of the reference's selection rule to PUCT.

    The reference's AlphaZero search maximises UCB1 (``node.py:41-42,75-88``);
    its own ``TreeNode.puct_value`` (``node.py:105-117``) divides by zero on an
    unvisited child.  The well-defined PUCT rule in the tree is
    ``SearchNode.puct_value`` (``rlzero/mcts/deepmind_mcts.py:149-151``):
    ``(n and W/n) + c*P*sqrt(Np)/(n+1)``.  This returns a context manager that
    installs that formula as ``TreeNode.uct_value`` and restores it on exit.
    """
    import contextlib
    import math

    @contextlib.contextmanager
    def _cm():
        old = ref.TreeNode.uct_value

        def puct(self, c_puct):
            n = self.explore_count
            return (n and self.total_reward / n) + c_puct * self.prior * math.sqrt(
                self._parent.explore_count) / (n + 1)

        ref.TreeNode.uct_value = puct
        try:
            yield
        finally:
            ref.TreeNode.uct_value = old

    return _cm()
