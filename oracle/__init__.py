"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the self-play MCTS hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker / the reported
CPU baseline.  The product (``rlzero_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Contents
--------
``ref_loader``   imports the live reference from ``/root/reference`` (only in
                 the authoring container) behind a 2-line ``gymnasium`` stub.
``pyoracle``     pure-Python restatement of the reference algorithm
                 (``rlzero/mcts/node.py``, ``rlzero/mcts/alphazero_mcts.py``,
                 ``rlzero/games/gomoku/gomoku_env.py``,
                 ``rlzero/games/gomoku/game.py``).
``c/``           plain-C restatement of the same search (board, win scan, UCT / PUCT selection,
                 expansion, sign-flipping backup, tree reuse, closed-form evaluators), OpenMP over
                 games, for parity at BASELINE size (8192 games x 800 playouts in seconds);
                 built by ``oracle/build_oracle.py`` into ``oracle/_build/`` (git-ignored).
``go_oracle``    Go rules (MiniGo / pettingzoo ``go_base`` restated) + the ``GoEnv`` wrapper; PARITY UNPINNED.
``dm_oracle``    the reference's second search driver ``DeepMindMCTS``, pinned by the live class.
``muzero_oracle`` the MuZero paper's search pseudocode; PARITY UNPINNED (no reference code).
``evaluators``   closed-form evaluators shared by both sides of a parity test.

Pinning: the reference's own tests hold no golden vector for this path
(SURVEY.md section 4), so the pin is (a) the live reference run in the
authoring container (``tests/test_oracle_vs_reference.py``) and (b) the
fixtures it generated, committed under ``tests/golden/`` together with
``scripts/make_golden.py``.
"""
