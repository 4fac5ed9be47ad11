"""Player shells of the reference API (rlzero/mcts/player.py:5-57)."""


class Player(object):
    """Base player: id / name accessors, abstract get_action / reset_player (player.py:5-30)."""

    def __init__(self, player_id=0, player_name=''):
        self.player_id = player_id
        self.player_name = player_name
        self.can_click = False

    def set_player_id(self, player_id):
        self.player_id = player_id

    def get_player_id(self):
        return self.player_id

    def get_player_name(self):
        return self.player_name

    def reset_player(self):
        raise NotImplementedError

    def get_action(self, game_env, **kwargs):
        raise NotImplementedError

    def __str__(self):
        return 'player'


class HumanPlayer(Player):
    """Reads 'row,col' from stdin until it names an empty square (player.py:33-57)."""

    def __init__(self, player_id=0, player_name=''):
        super().__init__(player_id, player_name)
        self.can_click = True

    def get_action(self, game_env, **kwargs):
        while True:
            move = -1
            try:
                text = input('Your move: ')
                move = game_env.location_to_move([int(tok, 10) for tok in text.split(',')])
            except Exception as exc:  # same forgiving behaviour as the reference
                print(exc)
            if move != -1 and move in game_env.leagel_actions():
                return move
            print('invalid move')

    def __str__(self):
        return 'HumanPlayer, id: {}, name {}.'.format(self.get_player_id(), self.get_player_name())
