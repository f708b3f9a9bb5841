"""Categorical -- torch mirror of svae/distributions/categorical.py:6-9."""
import torch

expectedstats = lambda natparam: torch.softmax(natparam, dim=-1)


def logZ(natparam):
    return torch.logsumexp(natparam, dim=-1).sum()
