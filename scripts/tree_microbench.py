#!/usr/bin/env python
"""Tree-kernel microbenchmark (GPU): waves of select -> closed-form eval -> expand+backup
at the bench workload's shape, timed per kernel with CUDA events.  Not the bench."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200.engine import ClosedFormEvaluator, SearchForest  # noqa: E402


def main():
    G = int(os.environ.get('G', 8192))
    H = int(os.environ.get('H', 15))
    n_playout = int(os.environ.get('NP', 800))
    f = SearchForest(G, H, 5, n_playout=n_playout, store_priors=bool(int(os.environ.get('PRI', 1))))
    print('HBM bytes: %.2f GB' % (f.hbm_bytes() / 1e9))
    ev = ClosedFormEvaluator(2)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    out = []
    for w in range(n_playout):
        e[0].record(); f.select(); e[1].record(); ev(f); e[2].record(); f.expand_backup(); e[3].record()
        if w in (0, 1, 100, 225, 226, 300, 500, 799):
            torch.cuda.synchronize()
            out.append((w, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))
    for r in out:
        print('wave %4d: select %.3f ms  eval %.3f ms  expand_backup %.3f ms' % r)
    f.raise_faults()
    # graph replay throughput
    f.reset_games()
    torch.cuda.synchronize()
    t0 = time.time()
    f.search(ev)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(json.dumps({'G': G, 'H': H, 'n_playout': n_playout, 'search_s': dt,
                      'sims_per_s_tree_only': G * n_playout / dt}))
    t0 = time.time(); f.root_policy(1.0); torch.cuda.synchronize(); t1 = time.time()
    f.advance(keep_subtree=True); torch.cuda.synchronize(); t2 = time.time()
    print('root_policy %.3f ms, advance %.3f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    f.raise_faults()


if __name__ == '__main__':
    main()
