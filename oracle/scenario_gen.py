"""TEST INFRASTRUCTURE -- numpy mirror of the device scenario generator (csrc/dcbf_gen.cuh, dcbf_gen_fields /
dcbf_gen_states).  Only tests/ may import it.

What is mirrored is the reference's generator /root/reference/rand_obs.py:31-81 (random_circle, random_obs,
gen_ran_obs_list) driven by a counter-based Philox4x32-10 stream instead of Python's global `random`, with the restart
rule that makes it terminate.  The reference seeds nothing, so it has no golden vectors for this path; parity is
  (1) this mirror == the CUDA kernel bit for bit on the obstacle fields (same draws, same IEEE operations),
  (2) the properties random_circle guarantees (ranges, two-decimal rounding, pairwise and keep-out separation), checked on
      both, and the same properties on fields drawn by the reference function itself when /root/reference is present.
Philox is pinned by the Random123 known-answer vectors (tests/test_scenario_gen_cpu.py).
"""
from __future__ import annotations

import math

import numpy as np

STREAM_FIELD, STREAM_MIX, STREAM_POS, STREAM_STATE = 1, 2, 3, 4
M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c, k0, k1):
    """c: uint32 array [..., 4] of counters, k0/k1 python ints -> uint32 array [..., 4]."""
    c = np.asarray(c, dtype=np.uint64)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = h1 ^ c1 ^ np.uint64(k0), l1, h0 ^ c3 ^ np.uint64(k1), l0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def uniform2(seed, stream, idx, blk):
    """two doubles in [0, 1) per (idx, blk) pair; idx / blk broadcast."""
    idx, blk = np.broadcast_arrays(np.asarray(idx, dtype=np.uint64), np.asarray(blk, dtype=np.uint64))
    c = np.stack([idx, blk, np.full_like(idx, stream), np.zeros_like(idx)], axis=-1)
    o = philox4x32_10(c, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF).astype(np.uint64)
    a = (o[..., 0] << np.uint64(32)) | o[..., 1]
    b = (o[..., 2] << np.uint64(32)) | o[..., 3]
    return (a >> np.uint64(11)).astype(np.float64) * 2.0 ** -53, (b >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def round2(x):
    return np.rint(np.float64(x) * 100.0) / 100.0


def _clear(x, y, r, o, half_gap):
    return (x - o[0]) ** 2 + (y - o[1]) ** 2 - (r + o[2] + 2 * half_gap) ** 2 >= 0


def make_field(seed, f, num, mix, margin=8.5, radius=1.0, half_gap=0.8, safe_dis=0.4, stall=2000, max_restarts=64):
    """One field -> (cir[Kc,3], elp[Ke,5], draws); scalar loop like rand_obs.py:31-54."""
    Kc, Ke = ((num + 1) // 2, num // 2) if mix else (num, 0)
    q, placed = 0, []
    for _ in range(max_restarts + 1):
        placed, tries = [], 0
        while tries < stall and len(placed) < num:
            # draw candidates in chunks of 64; only the ones the loop consumes count
            n = min(64, stall - tries)
            qs = q + np.arange(n)
            u0, u1 = uniform2(seed, STREAM_FIELD, f, 2 * qs)
            u2, _ = uniform2(seed, STREAM_FIELD, f, 2 * qs + 1)
            xs, ys, rs = round2(margin * u0), round2(margin * u1), round2((radius - 0.35) * u2 + 0.35)
            for i in range(n):
                x, y, r = float(xs[i]), float(ys[i]), float(rs[i])
                tries += 1
                q += 1
                if _clear(x, y, r, (10.0, 10.0, 0.3), half_gap) and _clear(x, y, r, (0.0, 0.0, 1.0), half_gap) \
                        and all(_clear(x, y, r, o, half_gap) for o in placed):
                    placed.append((x, y, r))
                    if len(placed) == num:
                        break
        if len(placed) == num:
            break
    if len(placed) != num:
        return np.full((Kc, 3), np.nan), np.full((Ke, 5), np.nan), -1
    cir, elp = np.zeros((Kc, 3)), np.zeros((Ke, 5))
    for i, (x, y, r) in enumerate(placed):
        if not mix or i % 2 == 0:
            cir[i // 2 if mix else i] = (x, y, r + safe_dis)
        else:
            u0, u1 = uniform2(seed, STREAM_MIX, f, i)
            ha = r * 0.5
            b = float(round2(ha * float(u0) + ha))
            phi = float(round2(math.floor(float(u1) * 181.0) * 3.141592653589793 / 180.0))
            elp[i // 2] = (x, y, r + safe_dis, b + safe_dis, phi)
    return cir, elp, q


def make_fields(seed, F, num, mix, **kw):
    out = [make_field(seed, f, num, mix, **kw) for f in range(F)]
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out]), np.array([o[2] for o in out], dtype=np.int32)


def clearance(cir, elp, x, y):
    """min level-set value over the (inflated) obstacles of one field."""
    h = np.inf
    for c in cir:
        h = min(h, (x - c[0]) ** 2 + (y - c[1]) ** 2 - c[2] * c[2])
    for e in elp:
        cp, sp = math.cos(e[4]), math.sin(e[4])
        a_ = (e[3] * cp) ** 2 + (e[2] * sp) ** 2
        b_ = 2.0 * cp * sp * (e[3] * e[3] - e[2] * e[2])
        c_ = (e[3] * sp) ** 2 + (e[2] * cp) ** 2
        dx, dy = x - e[0], y - e[1]
        h = min(h, a_ * dx * dx + b_ * dx * dy + c_ * dy * dy - (e[3] * e[2]) ** 2)
    return h


def make_states(seed, B, cir, elp, field=None, dd=False, goal=(10.0, 10.0), bvy_max=0.3, span=8.0, min_clear=0.05,
                jitter=0.3, max_attempts=64):
    """-> dict(x0, goal, leg, warm, last_u, attempts, margin) ; margin[b] = clearance - min_clear of the accepted position and
    of every rejected one (smallest absolute value), so a test can tell a genuine mismatch from a tie at the threshold."""
    field = np.zeros(B, dtype=np.int64) if field is None else np.asarray(field)
    nx = 3 if dd else 5
    x0, leg, att = np.full((B, nx), np.nan), np.zeros(B, dtype=np.int32), np.full(B, -1, dtype=np.int32)
    margin = np.full(B, np.inf)
    a_idx = np.arange(max_attempts)
    for b in range(B):
        u0, u1 = uniform2(seed, STREAM_POS, b, a_idx)
        px = py = math.nan
        for a in range(max_attempts):
            x, y = span * float(u0[a]), span * float(u1[a])
            h = clearance(cir[field[b]], elp[field[b]], x, y)
            margin[b] = min(margin[b], abs(h - min_clear))
            if h >= min_clear:
                px, py, att[b] = x, y, a + 1
                break
        s0, s1 = uniform2(seed, STREAM_STATE, b, 0)
        s2, s3 = uniform2(seed, STREAM_STATE, b, 1)
        lg = 1 if float(s1) < 0.5 else -1
        th = math.atan2(goal[1] - py, goal[0] - px) + (-jitter + (2.0 * jitter) * float(s0))
        leg[b] = lg
        if dd:
            x0[b] = (px, py, th)
        else:
            vbx = 0.4 + (0.8 - 0.4) * float(s2)
            vby = -lg * (0.15 + (bvy_max - 0.15) * float(s3))
            x0[b] = (px, py, math.cos(th) * vbx - math.sin(th) * vby, math.sin(th) * vbx + math.cos(th) * vby, th)
    g = np.tile(np.asarray(goal, dtype=np.float64), (B, 1))
    warm = np.tile(np.array([0.8, 0.0]), (B, 3)) if dd else np.tile(x0, (1, 3))
    last_u = np.tile(np.array([0.8, 0.0]), (B, 1)) if dd else None
    return dict(x0=x0, goal=g, leg=leg, warm=warm, last_u=last_u, attempts=att, margin=margin)
