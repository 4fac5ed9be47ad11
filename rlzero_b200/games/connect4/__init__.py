from .connect4_env import ConnectFourEnv

__all__ = ['ConnectFourEnv']
