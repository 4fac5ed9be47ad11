from .alphazero_mcts import AlphaZeroMCTS, AlphaZeroPlayer, softmax
from .node import TreeNode
from .player import HumanPlayer, Player

__all__ = ['AlphaZeroMCTS', 'AlphaZeroPlayer', 'TreeNode', 'Player', 'HumanPlayer', 'softmax']
