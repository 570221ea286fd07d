"""Initial values / parameters of the reference's IVP zoo (src/odecheckpts/ivps.py), as numpy."""

import numpy as np

THREE_BODY_MU = 0.012277471
THREE_BODY_T = 17.0652165601579625588917206249


def brusselator_u0(N):
    # ivps.py:147-153
    x0 = np.linspace(0, 1, num=N)
    return np.concatenate([np.sin(2 * np.pi * x0) + 1, 3.0 * np.ones(N)])[None, :]


def van_der_pol_u0():
    return np.array([[2.0], [0.0]])  # ivps.py:164-166


def rigid_body_u0():
    return np.array([[1.0, 0.0, 0.9]])  # diffeqzoo rigid_body


RIGID_BODY_PARAMS = (-2.0, 1.25, -0.5)


def three_body_u0():
    return np.array([[0.994, 0.0], [0.0, -2.00158510637908252240537862224]])


def pleiades_u0():
    # ivps.py:60-73
    x = np.array([3.0, 3.0, -1.0, -3.0, 2.0, -2.0, 2.0, 3.0, -3.0, 2.0, 0.0, 0.0, -4.0, 4.0])
    dx = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 1.75, -1.5, 0.0, 0.0, 0.0, -1.25, 1.0, 0.0, 0.0])
    return np.stack([x, dx])


def logistic_exact(t, u0=0.1):
    return u0 * np.exp(t) / (1.0 + u0 * (np.exp(t) - 1.0))
