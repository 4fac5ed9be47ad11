"""TEST INFRASTRUCTURE ONLY -- pure-Python MuZero search.  PARITY UNPINNED.

The reference (jianzhnie/RLZero) has no MuZero code.  This restates the pseudocode
published with the MuZero paper (Schrittwieser et al. 2020, supplementary ``pseudocode.py``):
``MinMaxStats``, ``Node``, ``run_mcts``, ``select_child``, ``ucb_score``, ``expand_node``,
``backpropagate``, ``add_exploration_noise``, in the two-player board-game setting (reward 0) with the
value-sign convention of the public re-implementations: a node's ``value_sum`` is from its own
``to_play`` view, a child's value enters the parent's score negated, and MinMaxStats is updated with
``reward + discount * -value``.  ``select_child`` takes ``max`` over ``(score, action, child)``
tuples, so ties go to the highest action.

Networks are passed in as callables so that a parity test can replay the device's network outputs:
``root_priors`` (dict action -> prior, already normalised over the legal actions and noise-mixed) and
``recurrent(sim_index, parent_node_id, action) -> (priors over the action space, value)``.
"""
import math


class MinMaxStats(object):

    def __init__(self, known_bounds=None):
        self.maximum = known_bounds[1] if known_bounds else -float('inf')
        self.minimum = known_bounds[0] if known_bounds else float('inf')

    def update(self, value):
        self.maximum = max(self.maximum, value)
        self.minimum = min(self.minimum, value)

    def normalize(self, value):
        if self.maximum > self.minimum:
            return (value - self.minimum) / (self.maximum - self.minimum)
        return value


class Node(object):

    def __init__(self, prior):
        self.visit_count = 0
        self.to_play = -1
        self.prior = prior
        self.value_sum = 0.0
        self.children = {}
        self.reward = 0
        self.node_id = -1           # creation order: 0 = root, i + 1 = created by simulation i

    def expanded(self):
        return len(self.children) > 0

    def value(self):
        if self.visit_count == 0:
            return 0
        return self.value_sum / self.visit_count


def ucb_score(cfg, parent, child, stats):
    pb_c = math.log((parent.visit_count + cfg.pb_c_base + 1) / cfg.pb_c_base) + cfg.pb_c_init
    pb_c *= math.sqrt(parent.visit_count) / (child.visit_count + 1)
    prior_score = pb_c * child.prior
    if child.visit_count > 0:
        value_score = stats.normalize(child.reward + cfg.discount * -child.value())
    else:
        value_score = 0
    return prior_score + value_score


def select_child(cfg, node, stats):
    best = None
    for action, child in node.children.items():
        key = (ucb_score(cfg, node, child, stats), action)
        if best is None or key > best[0]:
            best = (key, action, child)
    return best[1], best[2]


def run_mcts(cfg, root_to_play, root_priors, recurrent, n_actions):
    """Returns (root, stats, nodes in creation order, per-simulation (parent id, action))."""
    stats = MinMaxStats(cfg.known_bounds)
    root = Node(0)
    root.to_play = root_to_play
    root.node_id = 0
    for a in sorted(root_priors):
        root.children[a] = Node(root_priors[a])
    nodes = [root]
    trace = []
    for sim in range(cfg.num_simulations):
        node = root
        path = [node]
        to_play = root_to_play
        last_action = None
        while node.expanded():
            last_action, node = select_child(cfg, node, stats)
            to_play = 1 - to_play
            path.append(node)
        parent = path[-2]
        priors, value = recurrent(sim, parent.node_id, last_action)
        node.to_play = to_play
        node.node_id = len(nodes)
        nodes.append(node)
        trace.append((parent.node_id, last_action))
        for a in range(n_actions):
            node.children[a] = Node(priors[a])
        # backpropagate
        for nd in reversed(path):
            nd.value_sum += value if nd.to_play == to_play else -value
            nd.visit_count += 1
            stats.update(nd.reward + cfg.discount * -nd.value())
            value = nd.reward + cfg.discount * value
    return root, stats, nodes, trace
