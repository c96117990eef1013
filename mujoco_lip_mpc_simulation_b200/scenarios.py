"""Synthetic scenario batches for the BASELINE.json configs.

Obstacle fields restate the reference generator /root/reference/rand_obs.py:31-72 (rejection-sampled circles with
keep-out discs at the start and goal, every second obstacle turned into an ellipse in 'mix' mode) with a
termination guarantee: a field that makes no progress for `stall` draws is restarted.  All randomness comes from
numpy.random.default_rng(seed) (PCG64), so a (config, seed, B) triple names one exact batch on every machine.

Layouts (float64, C order):
    x0     [B,5]  LIP state (px, py, vx, vy, theta) in the map frame        (DD: [B,3] = (x, y, theta))
    goal   [B,2]
    leg    [B]    int32, +1/-1 (od_ev argument of solveMPCCBF)
    cir    [F,Kc,3] inflated circles (cx, cy, r+safe_dis)
    elp    [F,Ke,5] inflated ellipses (cx, cy, a+safe_dis, b+safe_dis, phi)
    field  [B]    int32 index of each scenario's obstacle field
    warm   [B,15] reference warm start u0 (DD: [B,6]); cold start = [x0,x0,x0] (MPC_LIP_sig_step.py:186-187)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class ScenarioBatch:
    form: str
    x0: np.ndarray
    goal: np.ndarray
    leg: np.ndarray
    cir: np.ndarray
    elp: np.ndarray
    field: np.ndarray
    warm: np.ndarray
    last_u: np.ndarray | None = None
    safe_dis: float = 0.4

    @property
    def B(self) -> int:
        return int(self.x0.shape[0])


def random_field(rng: np.random.Generator, num: int, margin: float = 8.5, radius: float = 1.0,
                 half_gap: float = 0.8, stall: int = 2000):
    """num circles [cx,cy,r]; pairwise centre distance >= r_i + r_j + 2*half_gap (rand_obs.py:31-54)."""
    while True:
        placed = [(10.0, 10.0, 0.3), (0.0, 0.0, 1.0)]
        tries = 0
        while len(placed) < num + 2 and tries < stall:
            tries += 1
            x = round(margin * rng.random(), 2)
            y = round(margin * rng.random(), 2)
            r = round((radius - 0.35) * rng.random() + 0.35, 2)
            if all((x - o[0]) ** 2 + (y - o[1]) ** 2 - (r + o[2] + 2 * half_gap) ** 2 >= 0 for o in placed):
                placed.append((x, y, r))
        if len(placed) == num + 2:
            return np.array(placed[2:], dtype=np.float64)


def split_mix(rng: np.random.Generator, circles: np.ndarray):
    """'mix' mode of rand_obs.py:57-72: even indices stay circles, odd ones become ellipses."""
    cir, elp = [], []
    for i, c in enumerate(circles):
        if i % 2 == 0:
            cir.append(c)
        else:
            a = c[2]
            b = round((a / 2) * rng.random() + (a / 2), 2)
            phi = round(int(rng.integers(0, 181)) * math.pi / 180, 2)
            elp.append([c[0], c[1], a, b, phi])
    return np.array(cir, dtype=np.float64).reshape(-1, 3), np.array(elp, dtype=np.float64).reshape(-1, 5)


def make_fields(seed: int, n_fields: int, num: int, mix: bool, half_gap: float, safe_dis: float):
    rng = np.random.default_rng(seed)
    cirs, elps = [], []
    for _ in range(n_fields):
        base = random_field(rng, num, half_gap=half_gap)
        if mix:
            c, e = split_mix(rng, base)
        else:
            c, e = base, np.zeros((0, 5))
        c = c + np.array([0, 0, safe_dis])
        if len(e):
            e = e + np.array([0, 0, safe_dis, safe_dis, 0])
        cirs.append(c)
        elps.append(e)
    return np.stack(cirs), np.stack(elps)


def _h_all(pos, cir, elp):
    """level-set values of all obstacles of one field at pos[..., 2] -> min over obstacles."""
    out = np.full(pos.shape[:-1], np.inf)
    for c in cir:
        out = np.minimum(out, (pos[..., 0] - c[0]) ** 2 + (pos[..., 1] - c[1]) ** 2 - c[2] ** 2)
    for e in elp:
        cp, sp = math.cos(e[4]), math.sin(e[4])
        a_ = (e[3] * cp) ** 2 + (e[2] * sp) ** 2
        b_ = 2 * cp * sp * (e[3] ** 2 - e[2] ** 2)
        c_ = (e[3] * sp) ** 2 + (e[2] * cp) ** 2
        dx, dy = pos[..., 0] - e[0], pos[..., 1] - e[1]
        out = np.minimum(out, a_ * dx * dx + b_ * dx * dy + c_ * dy * dy - (e[2] * e[3]) ** 2)
    return out


def make_batch(form: str, B: int, seed: int, n_fields: int | None = None, num_obs: int | None = None,
               goal=(10.0, 10.0), safe_dis: float = 0.4, bvy_max: float | None = None) -> ScenarioBatch:
    """Scenario distribution of SURVEY.md 8(d): config 2/5 = ('sig_step', K=6 circles), config 3 = ('modi',
    K=10 mixed), config 4 = ('dd', same fields as config 3)."""
    if form == "sig_step":
        num, mix, half_gap = num_obs or 6, False, 0.8
        bvy = bvy_max or 0.3
    elif form in ("modi", "dd"):
        num, mix, half_gap = num_obs or 10, True, 0.4
        bvy = bvy_max or 0.35
    else:
        raise ValueError(form)
    F = min(B, n_fields or 4096)
    cir, elp = make_fields(seed * 7919 + 13, F, num, mix, half_gap, safe_dis)
    rng = np.random.default_rng(seed)
    field = (np.arange(B) % F).astype(np.int32)
    pos = np.zeros((B, 2))
    todo = np.arange(B)
    while len(todo):
        cand = rng.random((len(todo), 2)) * 8.0
        ok = np.zeros(len(todo), dtype=bool)
        for k, b in enumerate(todo):
            f = field[b]
            ok[k] = _h_all(cand[k], cir[f], elp[f]) >= 0.05
        pos[todo[ok]] = cand[ok]
        todo = todo[~ok]
    g = np.broadcast_to(np.asarray(goal, dtype=np.float64), (B, 2)).copy()
    theta = np.arctan2(g[:, 1] - pos[:, 1], g[:, 0] - pos[:, 0]) + rng.uniform(-0.3, 0.3, B)
    leg = np.where(rng.random(B) < 0.5, 1, -1).astype(np.int32)
    if form == "dd":
        x0 = np.column_stack([pos, theta])
        warm = np.tile(np.array([0.8, 0.0]), (B, 3))
        last_u = np.tile(np.array([0.8, 0.0]), (B, 1))
        return ScenarioBatch(form, x0, g, leg, cir, elp, field, warm, last_u, safe_dis)
    vbx = rng.uniform(0.4, 0.8, B)
    vby = -leg * rng.uniform(0.15, bvy, B)
    vx = np.cos(theta) * vbx - np.sin(theta) * vby
    vy = np.sin(theta) * vbx + np.cos(theta) * vby
    x0 = np.column_stack([pos, vx, vy, theta])
    warm = np.tile(x0, (1, 3))
    return ScenarioBatch(form, x0, g, leg, cir, elp, field, warm, None, safe_dis)


def make_batch_device(solver, B: int, seed: int, n_fields: int | None = None, num_obs: int | None = None, goal=(10.0, 10.0),
                      safe_dis: float = 0.4):
    """The scenario distribution of make_batch drawn on the GPU (DcbfSolver.gen_fields / gen_states: dcbf_gen_fields,
    dcbf_gen_states), nothing touches the host: n_fields obstacle fields (default one per scenario, up to 65536) are built and
    installed as the solver's fields, then B start states on them.  Returns the dict of device tensors gen_states produces
    plus cir / elp / draws; pass x0, goal, leg, warm, field, last_u straight to DcbfSolver.solve."""
    form = {0: "sig_step", 1: "modi", 2: "dd"}[solver.form]
    num, mix, half_gap = (num_obs or 6, False, 0.8) if form == "sig_step" else (num_obs or 10, True, 0.4)
    F = min(B, n_fields or 65536)
    fields = solver.gen_fields(F, seed * 7919 + 13, num, mix=mix, half_gap=half_gap, safe_dis=safe_dis)
    import torch
    fld = (torch.arange(B, device=solver.tdev, dtype=torch.int32) % F).contiguous()
    out = solver.gen_states(B, seed, field=fld, goal=goal)
    out.update(fields)
    return out


def config1():
    """The reference's own single scenario, MPC_LIP_sig_step.py:553-568."""
    cir = (np.array([[1, 1, 0.5], [2, 2, 0.5], [6, 4, 0.8], [7, 7, 1.0]]) + np.array([0, 0, 0.32]))[None]
    x0 = np.array([[0.0, 0.0, 0.6, -0.3, 0.0]])
    return ScenarioBatch("sig_step", x0, np.array([[10.0, 10.0]]), np.array([1], dtype=np.int32), cir,
                         np.zeros((1, 0, 5)), np.zeros(1, dtype=np.int32), np.tile(x0, (1, 3)), None, 0.32)
