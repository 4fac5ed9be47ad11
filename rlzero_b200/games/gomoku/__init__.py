from .game import GameControl
from .gomoku_env import GomokuEnv, LeafEnvView

__all__ = ['GomokuEnv', 'GameControl', 'LeafEnvView']
