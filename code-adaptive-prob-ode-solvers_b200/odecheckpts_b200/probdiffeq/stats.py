"""``probdiffeq.stats``: the two calls of src/odecheckpts/ivpsolvers.py:80-81.

The backward marginalisation over the K checkpoints already ran on the device
(csrc/pn_smooth_kernel.cuh) when the solution was computed; these functions re-expose its
result in the shapes the reference indexes.
"""

from typing import NamedTuple


class Normal(NamedTuple):
    mean: object      # [..., n, d]
    cholesky: object  # [..., n, n] lower triangular


class MarkovSeq(NamedTuple):
    init: Normal          # marginals at checkpoints 1..K-1 (init.mean[-1] = terminal)
    marginals_all: Normal  # smoothed marginals at checkpoints 0..K-1


def markov_select_terminal(posterior):
    if not isinstance(posterior, MarkovSeq):
        raise TypeError("expected the .posterior of a solve_adaptive_save_at solution")
    return posterior


def markov_marginals(markov_seq, *, reverse):
    """Marginals of the checkpoint Markov sequence at t_0 .. t_{K-2} (the terminal marginal is
    ``markov_seq.init.mean[-1]``), as ivpsolvers.py:80-86 concatenates them."""
    if not reverse:
        raise NotImplementedError("only the backward (reverse=True) factorisation is produced by the solver")
    allm = markov_seq.marginals_all
    return Normal(allm.mean[:-1], allm.cholesky[:-1])
