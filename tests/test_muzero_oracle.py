"""CPU known-answer tests of the MuZero oracle (oracle/muzero_oracle.py).  The reference has no MuZero code, so the
oracle restates the pseudocode published with the paper; these vectors are worked out by hand from that pseudocode's
formulas (``ucb_score``, ``MinMaxStats``, the ``max`` over ``(score, action, child)`` tuples in ``select_child``,
``backpropagate`` in the two-player convention) and pin the restatement to them.  Parity for MuZero stays UNPINNED
against the reference."""
import math
from types import SimpleNamespace

from oracle import muzero_oracle as mz


def _cfg(n_sims, known_bounds=None):
    return SimpleNamespace(num_simulations=n_sims, discount=1.0, pb_c_base=19652, pb_c_init=1.25,
                           known_bounds=known_bounds)


def test_ucb_score_known_answer():
    """parent N = 10, child n = 2, prior 0.25, child value_sum -0.6 (value -0.3), discount 1, no bounds yet:
    pb_c = ln(19663 / 19652) + 1.25 = 1.2505595828710...; * sqrt(10) / 3 = 1.31820554387...; * 0.25 = 0.32955138596...
    value score = 0 + 1 * -(-0.3) = 0.3 (MinMaxStats does not normalise while max <= min)."""
    parent, child = mz.Node(0), mz.Node(0.25)
    parent.visit_count, child.visit_count, child.value_sum = 10, 2, -0.6
    s = mz.ucb_score(_cfg(1), parent, child, mz.MinMaxStats(None))
    assert abs(s - 0.6295513859685423) < 1e-15
    # an unvisited child scores its prior term only: pb_c * sqrt(10) / 1 * 0.25
    fresh = mz.Node(0.25)
    s0 = mz.ucb_score(_cfg(1), parent, fresh, mz.MinMaxStats(None))
    assert abs(s0 - 1.250559582871018 * math.sqrt(10) * 0.25) < 1e-15


def test_min_max_stats():
    st = mz.MinMaxStats(None)
    assert st.normalize(0.7) == 0.7                  # nothing seen yet: identity
    st.update(0.2)
    assert st.normalize(0.7) == 0.7                  # max == min: still identity
    st.update(-0.4)
    assert abs(st.normalize(0.0) - 0.4 / 0.6) < 1e-15 and st.normalize(0.2) == 1.0 and st.normalize(-0.4) == 0.0
    kb = mz.MinMaxStats((-1.0, 1.0))
    assert kb.normalize(0.0) == 0.5


def test_three_simulations_by_hand():
    """Two actions, priors 0.5 / 0.5 everywhere, every network value 0.
    sim 0: the root has N = 0, so both prior terms are pb_c * sqrt(0) = 0: a tie, and max over (score, action) takes the
           HIGHEST action, 1.  Node 1 = child 1.
    sim 1: root N = 1: child 1 (n = 1) scores pb_c * 1 / 2 * 0.5 + 0, child 0 (n = 0) pb_c * 1 / 1 * 0.5: action 0.  Node 2.
    sim 2: root N = 2, both children n = 1, equal priors and values: tie -> action 1; inside node 1 (N = 1) both
           grandchildren are unvisited with equal priors: tie -> action 1.  Node 3 = grandchild (1, 1).
    Visit counts: root 3, child 0: 1, child 1: 2, grandchild (1,1): 1; all value sums 0."""
    calls = []

    def recurrent(sim, parent_id, action):
        calls.append((sim, parent_id, action))
        return [0.5, 0.5], 0.0

    root, stats, nodes, trace = mz.run_mcts(_cfg(3), 0, {0: 0.5, 1: 0.5}, recurrent, 2)
    assert trace == [(0, 1), (0, 0), (1, 1)] and calls == [(0, 0, 1), (1, 0, 0), (2, 1, 1)]
    assert root.visit_count == 3
    assert root.children[0].visit_count == 1 and root.children[1].visit_count == 2
    assert root.children[1].children[1].visit_count == 1 and root.children[1].children[0].visit_count == 0
    assert [nd.node_id for nd in nodes] == [0, 1, 2, 3]
    assert nodes[1] is root.children[1] and nodes[2] is root.children[0] and nodes[3] is root.children[1].children[1]
    assert all(nd.value_sum == 0.0 for nd in nodes)
    assert (nodes[1].to_play, nodes[3].to_play) == (1, 0)         # players alternate below the root (to_play 0)


def test_backpropagation_signs_two_player():
    """One simulation whose leaf value is +0.5 for the player to move AT THE LEAF (player 1, the root being player 0):
    the leaf's value_sum gets +0.5, the root's -0.5 (the other player's view); MinMaxStats sees discount * -value of
    each node: -0.5 for the leaf and +0.5 for the root."""
    root, stats, nodes, _ = mz.run_mcts(_cfg(1), 0, {0: 0.5, 1: 0.5}, lambda s, p, a: ([0.5, 0.5], 0.5), 2)
    leaf = nodes[1]
    assert leaf.to_play == 1 and leaf.value_sum == 0.5 and root.value_sum == -0.5
    assert (stats.minimum, stats.maximum) == (-0.5, 0.5)
