"""Dirichlet -- torch mirror of svae/distributions/dirichlet.py:5-11."""
import torch


def expectedstats(natparam):
    alpha = natparam + 1
    return torch.digamma(alpha) - torch.digamma(alpha.sum(-1, keepdim=True))


def logZ(natparam):
    alpha = natparam + 1
    return (torch.lgamma(alpha).sum(-1) - torch.lgamma(alpha.sum(-1))).sum()
