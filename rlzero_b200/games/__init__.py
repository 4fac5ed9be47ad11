from .base_env import BaseEnv
from .gomoku.game import GameControl
from .gomoku.gomoku_env import GomokuEnv

__all__ = ['BaseEnv', 'GomokuEnv', 'GameControl']
