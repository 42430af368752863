"""Shared helpers for the parity tests (oracle is the checker, never the product)."""
import json
import os
import numpy as np
import torch
from oracle import maskedsst_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a)).double().flatten() if not torch.is_tensor(a) else a.detach().double().cpu().flatten()
    b = torch.as_tensor(np.asarray(b)).double().flatten() if not torch.is_tensor(b) else b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def hash_str(s):
    h = 0
    for ch in s:
        h = (h * 131 + ord(ch)) % 2147483647
    return h


def proj_vec(shape, key):
    rng = np.random.Generator(np.random.PCG64(abs(hash_str(key)) % (2 ** 31)))
    return rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=shape)


def grad_rows(named):
    """Same (l2, sum, signed projection) summary as tests/golden/make_golden.py."""
    out = {}
    for k, g in named:
        g = g.detach().double().cpu().numpy()
        out[k] = np.array([np.sqrt((g * g).sum()), g.sum(), (g * proj_vec(g.shape, k)).sum()])
    return out


def check_grad_rows(got, names, rows, tol, floor=1e-7):
    """Every tensor's (l2 norm, sum, +-1 projection) must match the golden triple.  With e = got - gold
    and ||e|| <= tol*||gold||:  |d norm| <= ||e||;  |d proj| ~ ||e|| for a random sign vector (8x slack);
    |d sum| <= sqrt(n)||e|| worst case (64x slack covers n <= 4096 worst-case, far more typically)."""
    worst = 0.0
    for k, r in zip(names, rows):
        k = str(k)
        assert k in got, f"missing grad for {k}"
        scale = max(r[0], floor)
        e = [abs(got[k][0] - r[0]) / scale, abs(got[k][2] - r[2]) / (8 * scale), abs(got[k][1] - r[1]) / (64 * scale)]
        worst = max(worst, *e)
        assert max(e) <= tol, f"{k}: got {got[k]} vs golden {r} (scaled errs {e}, tol {tol})"
    return worst
