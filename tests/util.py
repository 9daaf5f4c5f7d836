"""Shared helpers for the parity tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

FWD_TOL = 1e-4     # forward bar from BASELINE.json north_star: 1e-4 relative fp32


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def max_rel(a, b):
    """max|a-b| / max|b| -- the forward parity measure of SURVEY.md section 8(c)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def l2_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def sample(t, cap=1024):
    """Same sub-sampling as oracle/make_golden.py::sample."""
    t = t.detach().reshape(-1)
    if t.numel() <= cap:
        return t.clone()
    step = t.numel() // cap
    return t[::step][:cap].clone()
