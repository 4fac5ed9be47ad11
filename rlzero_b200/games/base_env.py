"""Abstract env interface of the reference (rlzero/games/base_env.py:7-33), without the
``gymnasium.Env`` base class (gymnasium is not a dependency of this package)."""
import copy


class BaseEnv(object):

    def render(self):
        raise NotImplementedError

    def current_player(self):
        raise NotImplementedError

    def legal_actions(self, player):
        raise NotImplementedError

    def returns(self):
        raise NotImplementedError

    def clone(self):
        return copy.deepcopy(self)

    def is_terminal(self):
        raise NotImplementedError
