"""Read-only ``TreeNode`` view over a search tree that lives in HBM.

The reference keeps one Python object per child (rlzero/mcts/node.py:7-30).  Here the tree
is a structure of arrays on the device (see engine.SearchForest); this class presents a
host snapshot of it with the reference's attribute names so that code (and tests) written
against ``mcts._root._children[a].explore_count`` keep working.
"""


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
        self._kids = None

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

    def is_leaf(self):
        return self._children == {}

    def is_root(self):
        return self._parent is None

    def __str__(self):
        return 'TreeNode: {MCTSNode, Total Value:  %s, Num Visits: %s}' % (
            self.total_reward, self.explore_count)
