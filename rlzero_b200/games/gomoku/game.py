"""``GameControl`` with the reference's API (rlzero/games/gomoku/game.py:12-137): match runner
and self-play episode recorder for ONE game driven from Python.  The batched equivalent that
keeps whole generations of episodes on the device is ``rlzero_b200.selfplay``."""
import numpy as np

from .gomoku_env import Error


class GameControl(object):
    """game server."""

    def __init__(self, game_env):
        self.game_env = game_env
        self.visualTool = None

    def set_player_symbol(self, start_player):
        first = self.game_env.players[start_player] == self.game_env.players[0]
        self.player1_symbol, self.player2_symbol = ('X', 'O') if first else ('O', 'X')

    def graphic(self, game_env, player1, player2):
        """Draw the board and show game info (game.py:29-59)."""
        n = game_env.board_size
        id1 = player1 if isinstance(player1, int) else player1.get_player_id()
        id2 = player2 if isinstance(player2, int) else player2.get_player_id()
        print('Player', player1, self.player1_symbol.rjust(3))
        print('Player', player2, self.player2_symbol.rjust(3))
        print()
        for x in range(n):
            print('{0:8}'.format(x), end='')
        print('\r\n')
        for i in range(n - 1, -1, -1):
            print('{0:4d}'.format(i), end='')
            for j in range(n):
                p = game_env.states.get(i * n + j, -1)
                sym = self.player1_symbol if p == id1 else self.player2_symbol if p == id2 else '_'
                print(sym.center(8), end='')
            print('\r\n\r\n')

    def start_play(self, player1, player2, start_player=0, is_shown=True):
        """start a game between two players (game.py:61-94); returns the winner id or -1.
        As in the reference, ``start_player`` only picks the symbols: the env is reset() with
        its default, so player id 0 (= player1) always moves first."""
        if start_player not in (0, 1):
            raise Error(f'{start_player} should be 0 (player1 first) or 1 (player2 first)')
        self.game_env.reset()
        p1, p2 = self.game_env.players
        player1.set_player_id(p1)
        player2.set_player_id(p2)
        self.set_player_symbol(start_player)
        seats = {p1: player1, p2: player2}
        if is_shown:
            self.graphic(self.game_env, player1, player2)
        while True:
            mover = seats[self.game_env.current_player()]
            self.game_env.step(mover.get_action(self.game_env))
            if is_shown:
                self.graphic(self.game_env, player1, player2)
            end, winner = self.game_env.game_end_winner()
            if end:
                if is_shown:
                    print('Game end. Winner is', seats[winner]) if winner != -1 else print('Game end. Tie')
                return winner

    def start_self_play(self, player, is_shown=False, temperature=1e-3):
        """One self-play episode (game.py:96-134): returns ``(winner, zip(states, mcts_probs,
        winners_z))`` with z = +1 for plies of the winner, -1 for the loser's, 0 on a tie."""
        self.game_env.reset()
        p1, p2 = self.game_env.players
        states, mcts_probs, movers = [], [], []
        self.set_player_symbol(start_player=0)
        while True:
            move, move_probs = player.get_action(self.game_env, temperature=temperature,
                                                 return_prob=True)
            states.append(self.game_env.current_state())
            mcts_probs.append(move_probs)
            movers.append(self.game_env.current_player())
            self.game_env.step(move)
            if is_shown:
                self.graphic(self.game_env, p1, p2)
            end, winner = self.game_env.game_end_winner()
            if end:
                winners_z = np.zeros(len(movers))
                if winner != -1:
                    movers = np.array(movers)
                    winners_z[movers == winner] = 1.0
                    winners_z[movers != winner] = -1.0
                player.reset_player()
                if is_shown:
                    print('Game end. Winner is player:', winner) if winner != -1 else print('Game end. Tie')
                return winner, zip(states, mcts_probs, winners_z)

    def __str__(self):
        return 'Game'
