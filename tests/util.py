"""Shared test helpers."""
import numpy as np
import torch


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def load_state_into(module, state):
    module.load_state_dict({k: v.clone() for k, v in state.items()})
    return module


def clone_state(state):
    return type(state)((k, v.detach().clone()) for k, v in state.items())
