"""``TreeNode`` with the reference's attributes and methods (rlzero/mcts/node.py:7-184).

The search trees live in HBM as a structure of arrays (engine.SearchForest); the kernels never
touch Python objects.  This class serves the code that does: a node built with
``TreeNode(parent, prior)`` is a plain host node with the reference's behaviour (``expand``,
``select``, ``update_recursive`` ... work on it as they do there), and ``from_snapshot`` presents
a host copy of a device tree through the same interface, children materialised lazily, so that
``mcts._root._children[a].explore_count`` -- and ``mcts._root.select(c_puct)`` -- keep working.
Mutating a snapshot changes the host copy only; the device tree is advanced by the search.
"""
import math

import numpy as np


class TreeNode(object):
    __slots__ = ('_snap', '_node', '_parent', '_action', 'explore_count', 'total_reward', 'prior',
                 '_kids')

    def __init__(self, parent=None, prior=1.0, _snap=None, _node=-1, _action=-1, _n=0, _w=0):
        self._snap = _snap
        self._node = _node        # index of this node's edge block, -1 if unexpanded
        self._parent = parent
        self._action = _action
        self.explore_count = _n
        self.total_reward = _w
        self.prior = prior
        self._kids = None if _snap is not None else {}

    @classmethod
    def from_snapshot(cls, snap):
        """Root view of ``SearchForest.dump_tree(g)``."""
        return cls(None, 1.0, _snap=snap, _node=0 if snap['n_nodes'] > 0 else -1,
                   _n=snap['root_N'], _w=snap['root_W'] if snap['root_N'] else 0)

    @property
    def _children(self):
        if self._kids is None:
            kids = {}
            if self._snap is not None and self._node >= 0:
                s, i = self._snap, self._node
                for a in range(s['N'].shape[1]):
                    n = int(s['N'][i, a])
                    if n < 0:
                        continue
                    ch = int(s['child'][i, a]) if n >= 1 else -1
                    kids[a] = TreeNode(self, float(s['P'][i, a]), _snap=s, _node=ch if ch >= 0 else -1,
                                       _action=a, _n=n, _w=float(s['W'][i, a]) if n >= 1 else 0)
            self._kids = kids
        return self._kids

    @property
    def children(self):
        return self._children

    @property
    def parent(self):
        return self._parent

    # ------------------------------------------------------------------ scores
    def uct_value(self, c_puct):
        """UCB1 score of this child (node.py:75-88): +inf while the parent or the child is unvisited."""
        n_parent, n = self._parent.explore_count, self.explore_count
        if n_parent == 0 or n == 0:
            return float('inf')
        return self.total_reward / n + c_puct * math.sqrt(math.log(n_parent) / n)

    ucb_value = uct_value         # node.py:90-103 is the same rule under a second name

    def puct_value(self, c_puct):
        """node.py:105-117; like the reference it divides by explore_count, so an unvisited child raises
        ZeroDivisionError (the search itself never calls it; the PUCT mode of the kernels is
        ``SearchNode.puct_value``, deepmind_mcts.py:149-151)."""
        u = self.prior * math.sqrt(self._parent.explore_count) / (self.explore_count + 1)
        return self.total_reward / self.explore_count + c_puct * u

    # ------------------------------------------------------------- tree policy
    def select(self, c_puct):
        """(action, child) of the highest ``uct_value``; the first maximum wins (node.py:32-42)."""
        kids = self._children
        if not kids:
            raise ValueError('Node has no children.')
        return max(kids.items(), key=lambda item: item[1].uct_value(c_puct))

    def expand(self, action_priors, add_noise=False):
        """One child per (action, prior) not present yet; with ``add_noise`` every prior becomes
        0.75 p + 0.25 Dir(0.3) drawn from the global numpy stream (node.py:44-73)."""
        pairs = list(action_priors)
        kids = self._children
        noise = np.random.dirichlet(0.3 * np.ones(len(pairs))) if add_noise else None
        for i, (action, prob) in enumerate(pairs):
            if action in kids:
                continue
            kids[action] = TreeNode(self, prob if noise is None else 0.75 * prob + 0.25 * noise[i])

    def update(self, value):
        """One more visit with ``value`` (node.py:119-133)."""
        self.explore_count += 1
        self.total_reward += value

    def update_recursive(self, leaf_value):
        """``update`` on every ancestor first, the sign alternating per level (node.py:135-144)."""
        if self._parent:
            self._parent.update_recursive(-leaf_value)
        self.update(leaf_value)

    def is_leaf(self):
        return self._children == {}

    def is_root(self):
        return self._parent is None

    def __str__(self):
        return 'TreeNode: {MCTSNode, Total Value:  %s, Num Visits: %s}' % (
            self.total_reward, self.explore_count)
