from .alphazero_mcts import AlphaZeroMCTS, AlphaZeroPlayer, softmax
from .deepmind_mcts import DeepMindMCTS, Evaluator, MCTSBot, RandomRolloutEvaluator, SearchNode
from .node import TreeNode
from .player import HumanPlayer, Player
from .rollout_mcts import RolloutMCTS, RolloutPlayer

__all__ = ['AlphaZeroMCTS', 'AlphaZeroPlayer', 'RolloutMCTS', 'RolloutPlayer', 'TreeNode', 'Player',
           'HumanPlayer', 'softmax', 'DeepMindMCTS', 'MCTSBot', 'SearchNode', 'Evaluator',
           'RandomRolloutEvaluator']
