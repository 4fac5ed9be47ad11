from .go_env import BaseEnvTimestep, GoBoards, GoEnv

__all__ = ['GoEnv', 'GoBoards', 'BaseEnvTimestep']
