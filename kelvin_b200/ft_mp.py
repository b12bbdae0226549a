"""Zeroth-order grand potential sums (kelvin/ft_mp.py:12-19)."""


def mp0(g0):
    return g0.sum()


def ump0(g0a, g0b):
    return g0a.sum() + g0b.sum()
