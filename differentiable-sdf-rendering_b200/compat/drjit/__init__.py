"""`drjit` as the reference's drivers use it (python/shape_opt.py:60,83,90-98): gradient scoping and reverse-mode AD on torch tensors."""
import math

import torch

pi = math.pi


def suspend_grad():
    return torch.no_grad()


def backward(x):
    if torch.is_tensor(x) and x.requires_grad:
        x.sum().backward()


def grad_enabled(x):
    return bool(torch.is_tensor(x) and x.requires_grad)


def detach(x, preserve_type=True):
    return x.detach() if torch.is_tensor(x) else x


def prod(x):
    r = 1
    for v in x:
        r *= int(v)
    return r


def grad(x):
    return x.grad


def enable_grad(x):
    x.requires_grad_(True)
