"""TEST INFRASTRUCTURE ONLY -- closed-form evaluators for parity tests.

A parity test must hand "the same network outputs" to the reference search and
to the CUDA search (BASELINE.json north_star).  These evaluators are pure
functions of the board whose outputs are dyadic rationals, so every sum the
search forms is exact in fp64 and there is no rounding to disagree about.  The
CUDA library computes the same formulas on the device
(``rlzero_b200/csrc/rz_eval.cuh``, ids below); here they are written against
the duck-typed env the reference's ``policy_value_fn`` receives
(``alphazero_mcts.py:27-31,59``): ``leagel_actions()``, ``states``,
``last_move``.

EVAL_KAT   value = ((17*stones + 31*(last_move+1)) % 13 - 6) / 8, uniform
           prior float32(1/len(legal)) -- the evaluator SURVEY.md section 4
           used for KAT A-F.
EVAL_HASH  value and priors depend on every stone (below).
EVAL_ZERO  value 0, uniform prior.
"""
import numpy as np

EVAL_ZERO = 0
EVAL_KAT = 1
EVAL_HASH = 2

_M32 = 0xFFFFFFFF


def board_hash(states, last_move):
    h = 0
    for m, p in states.items():
        h += (m + 1) * (m + 1) * (3 + 4 * p)
    h += 7 * (last_move + 1)
    h &= _M32
    return (h * 2654435761) & _M32


def value_of(eval_id, states, last_move):
    if eval_id == EVAL_ZERO:
        return 0.0
    if eval_id == EVAL_KAT:
        return ((17 * len(states) + 31 * (last_move + 1)) % 13 - 6) / 8.0
    if eval_id == EVAL_HASH:
        h = board_hash(states, last_move)
        return (((h >> 16) % 129) - 64) / 64.0
    raise ValueError(eval_id)


def priors_of(eval_id, states, last_move, legal):
    """float64 array (values exactly representable in float32), one per legal move."""
    if len(legal) == 0:
        return np.zeros(0, dtype=np.float64)
    if eval_id in (EVAL_ZERO, EVAL_KAT):
        p = np.float32(1.0) / np.float32(len(legal))
        return np.full(len(legal), float(p), dtype=np.float64)
    if eval_id == EVAL_HASH:
        h = board_hash(states, last_move)
        a = np.asarray(legal, dtype=np.int64)
        return (((a * 29 + (h >> 8)) % 32) + 1) / 256.0
    raise ValueError(eval_id)


def make_policy_value_fn(eval_id):
    """A ``policy_value_fn(env)`` for the reference / the restatement."""

    def fn(env):
        legal = list(env.leagel_actions())
        pri = priors_of(eval_id, env.states, env.last_move, legal)
        return zip(legal, [float(x) for x in pri]), value_of(
            eval_id, env.states, env.last_move)

    fn.eval_id = eval_id
    return fn
