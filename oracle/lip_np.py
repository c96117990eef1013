"""ORACLE (test infrastructure, not product code) -- numpy restatement of the reference NLP callbacks.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

What is restated (reference paths are relative to /root/reference):
  * LIP step-to-step model constants        MPC_LIP_sig_step.py:47-86  (A, B, W, M_A, M_B, dx_du, dP_du)
  * sig_step callbacks                      MPC_LIP_sig_step.py:372-548 (objective, gradient, constraints, jacobian)
  * modi callbacks (circles+ellipses+f_en)  MPC_LIP_modi.py:430-655
  * DD (unicycle) callbacks                 MPC_DD_sig_step.py:351-572
  * constraint-bound vectors, goal shift,
    obstacle selection, warm-start rule     MPC_LIP_sig_step.py:184-253, MPC_LIP_modi.py:197-271,325-338,
                                            MPC_DD_sig_step.py:123-141

All callbacks work in the reference's own decision space (u in R^15 for the LIP variants, R^6 for DD) and
return plain float64 ndarrays with the reference's row ordering.  Nothing here is vectorised over problems:
it is the slow, literal checker.

PARITY PIN: tests/golden/callbacks_*.npz were produced by running the *reference's own* LIP_Prob classes
(oracle/gen_golden.py, in the build container where /root/reference exists) and this file is checked
against them in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------------------------------------
# constants of the LIP model (MPC_LIP_sig_step.py:17-20,47-65)
# --------------------------------------------------------------------------------------------------------
HEIGHT = 1.0
GRAV = 9.81
BETA = math.sqrt(GRAV / HEIGHT)
DT = 0.4
N_STEPS = 3


@dataclass(frozen=True)
class Formulation:
    """Weights and limits that differ between the three reference files."""
    name: str
    p: float          # extra terminal-ish weight on step 1 position
    q: float          # position weight
    r: float          # heading weight
    gamma: float      # D-CBF decay
    s_turn: float     # speed/turn coupling coefficient
    has_fen: bool     # modi: extra row s*|dtheta| + v_bx
    bvx_min: float = 0.4
    bvx_max: float = 0.8
    bvy_min: float = 0.15
    bvy_max: float = 0.3
    leg_sq: float = 0.09
    ang_max: float = math.pi / 16
    t_smooth: float = 0.0  # DD only: control smoothness weight
    max_iter: int = 20


# MPC_LIP_sig_step.py:34-40,340-353,269
SIG_STEP = Formulation("sig_step", p=2.0, q=1.0, r=15.0, gamma=0.4, s_turn=0.014 * 180 / math.pi,
                       has_fen=False, bvy_max=0.3, max_iter=20)
# MPC_LIP_modi.py:35-41,397-411,287
MODI = Formulation("modi", p=0.0, q=1.0, r=50.0, gamma=0.2, s_turn=0.024 * 180 / math.pi,
                   has_fen=True, bvy_max=0.35, max_iter=30)
# MPC_DD_sig_step.py:33-37,323-338,183
DD = Formulation("dd", p=0.0, q=1.0, r=50.0, gamma=0.2, s_turn=0.024 * 180 / math.pi,
                 has_fen=True, t_smooth=2.0, max_iter=40)


@dataclass
class LipModel:
    """Matrices built once per planner object (MPC_LIP_sig_step.py:47-86)."""
    A: np.ndarray = field(init=False)
    B: np.ndarray = field(init=False)
    W: np.ndarray = field(init=False)
    M_A: np.ndarray = field(init=False)
    M_B: np.ndarray = field(init=False)
    dx_du: np.ndarray = field(init=False)   # 20 x 15
    dP_du: np.ndarray = field(init=False)   #  9 x 15

    def __post_init__(self):
        ch = math.cosh(BETA * DT)
        sh = math.sinh(BETA * DT)
        A = np.eye(5)
        A[0, 0] = A[1, 1] = A[2, 2] = A[3, 3] = ch
        A[0, 2] = A[1, 3] = sh / BETA
        A[2, 0] = A[3, 1] = sh * BETA
        B = np.zeros((5, 3))
        B[0, 0] = B[1, 1] = 1.0 - ch
        B[2, 0] = B[3, 1] = -sh * BETA
        B[4, 2] = 1.0
        wa, wb = 5.0, 1.0
        den = wa * (ch - 1.0) ** 2 + wb * (sh * BETA) ** 2
        c_h = -wa * (ch - 1.0) / den
        s_h = -wb * sh * BETA / den
        W = np.zeros((3, 5))
        W[0, 0] = W[1, 1] = c_h
        W[0, 2] = W[1, 3] = s_h
        W[2, 4] = 1.0
        self.A, self.B, self.W = A, B, W
        self.M_A = A - B @ W @ A
        self.M_B = B @ W
        blocks = [self.M_B, self.M_A @ self.M_B, self.M_A @ self.M_A @ self.M_B]
        dx = np.zeros((20, 15))
        dp = np.zeros((9, 15))
        pl = [W, -W @ A @ self.M_B, -W @ A @ self.M_A @ self.M_B]
        for row in range(1, 4):          # state index
            for col in range(row):       # control index
                dx[5 * row:5 * row + 5, 5 * col:5 * col + 5] = blocks[row - 1 - col]
        for row in range(3):
            for col in range(row + 1):
                dp[3 * row:3 * row + 3, 5 * col:5 * col + 5] = pl[row - col]
        self.dx_du, self.dP_du = dx, dp


_MODEL = None


def model() -> LipModel:
    global _MODEL
    if _MODEL is None:
        _MODEL = LipModel()
    return _MODEL


# --------------------------------------------------------------------------------------------------------
# obstacle level-set functions (MPC_LIP_sig_step.py:499-510, MPC_LIP_modi.py:586-617)
# --------------------------------------------------------------------------------------------------------
def h_circle(c, x, y):
    return (x - c[0]) ** 2 + (y - c[1]) ** 2 - c[2] ** 2


def dh_circle(c, x, y):
    return 2.0 * (x - c[0]), 2.0 * (y - c[1])


def ellipse_coeffs(e):
    """e = [cx, cy, a, b, phi] -> (a', b', c', rhs) of  a'dx^2 + b'dxdy + c'dy^2 - rhs."""
    cp, sp = math.cos(e[4]), math.sin(e[4])
    a_ = (e[3] * cp) ** 2 + (e[2] * sp) ** 2
    b_ = 2.0 * cp * sp * (e[3] ** 2 - e[2] ** 2)
    c_ = (e[3] * sp) ** 2 + (e[2] * cp) ** 2
    return a_, b_, c_, (e[3] * e[2]) ** 2


def h_ellipse(e, x, y):
    a_, b_, c_, rhs = ellipse_coeffs(e)
    dx, dy = x - e[0], y - e[1]
    return a_ * dx * dx + b_ * dx * dy + c_ * dy * dy - rhs


def dh_ellipse(e, x, y):
    a_, b_, c_, _ = ellipse_coeffs(e)
    dx, dy = x - e[0], y - e[1]
    return 2.0 * a_ * dx + b_ * dy, 2.0 * c_ * dy + b_ * dx


# --------------------------------------------------------------------------------------------------------
# LIP callbacks (u in R^15)
# --------------------------------------------------------------------------------------------------------
def lip_rollout(xk, u):
    """x[0..3] (4x5) and p[0..2] (3x3) from u (MPC_LIP_sig_step.py:376-381)."""
    m = model()
    x = np.zeros((4, 5))
    p = np.zeros((3, 3))
    x[0] = np.asarray(xk, dtype=np.float64).ravel()
    u = np.asarray(u, dtype=np.float64).ravel()
    for i in range(3):
        ui = u[5 * i:5 * i + 5]
        p[i] = m.W @ (ui - m.A @ x[i])
        x[i + 1] = m.M_A @ x[i] + m.M_B @ ui
    return x, p


def lip_objective(form: Formulation, xk, goal, u):
    x, _ = lip_rollout(xk, u)
    g = np.asarray(goal, dtype=np.float64).ravel()
    cost = 0.0
    for i in range(1, 4):
        d = x[i, 0:2] - g
        tar = math.atan2(g[1] - x[i, 1], g[0] - x[i, 0])
        cost += form.q * (d @ d) + form.r * (x[i, 4] - tar) ** 2
    d1 = x[1, 0:2] - g
    cost += form.p * (d1 @ d1)
    return float(cost)


def lip_gradient(form: Formulation, xk, goal, u):
    m = model()
    x, _ = lip_rollout(xk, u)
    g = np.asarray(goal, dtype=np.float64).ravel()
    out = np.zeros(15)
    for i in range(1, 4):
        w = form.q + (form.p if i == 1 else 0.0)
        out += 2.0 * w * ((x[i, 0] - g[0]) * m.dx_du[5 * i] + (x[i, 1] - g[1]) * m.dx_du[5 * i + 1])
        dx_, dy_ = g[0] - x[i, 0], g[1] - x[i, 1]
        tar = math.atan2(dy_, dx_)
        dtar = (dx_ * (-m.dx_du[5 * i + 1]) - dy_ * (-m.dx_du[5 * i])) / (dx_ * dx_ + dy_ * dy_)
        out += 2.0 * form.r * (x[i, 4] - tar) * (m.dx_du[5 * i + 4] - dtar)
    return out


def lip_bounds(form: Formulation, leg: int, n_cir: int, n_elp: int = 0):
    """cl, cu in reference row order (MPC_LIP_sig_step.py:193-227, MPC_LIP_modi.py:203-245)."""
    cl, cu = [], []
    for i in range(3):
        plus = (leg > 0) == (i % 2 == 0)
        lo, hi = (form.bvy_min, form.bvy_max) if plus else (-form.bvy_max, -form.bvy_min)
        cl += [form.bvx_min, lo] + [0.0] * (n_cir + n_elp) + [0.0, -form.ang_max]
        cu += [form.bvx_max, hi] + [np.inf] * (n_cir + n_elp) + [form.leg_sq, form.ang_max]
        if form.has_fen:
            cl.append(form.bvx_min)
            cu.append(form.bvx_max)
    return np.array(cl), np.array(cu)


def lip_constraints(form: Formulation, xk, circles, ellipses, u):
    x, p = lip_rollout(xk, u)
    rows = []
    for i in range(3):
        th = x[i + 1, 4]
        c, s = math.cos(th), math.sin(th)
        vbx = c * x[i + 1, 2] + s * x[i + 1, 3]
        vby = -s * x[i + 1, 2] + c * x[i + 1, 3]
        rows += [vbx, vby]
        for cir in circles:
            rows.append(h_circle(cir, x[i + 1, 0], x[i + 1, 1]) + (form.gamma - 1.0) * h_circle(cir, x[i, 0], x[i, 1]))
        for elp in ellipses:
            rows.append(h_ellipse(elp, x[i + 1, 0], x[i + 1, 1]) + (form.gamma - 1.0) * h_ellipse(elp, x[i, 0], x[i, 1]))
        rows.append((x[i, 0] - p[i, 0]) ** 2 + (x[i, 1] - p[i, 1]) ** 2)
        rows.append(p[i, 2])
        if form.has_fen:
            rows.append(form.s_turn * abs(p[i, 2]) + vbx)
    return np.array(rows, dtype=np.float64)


def lip_jacobian(form: Formulation, xk, circles, ellipses, u):
    m = model()
    x, p = lip_rollout(xk, u)
    rows = []
    for i in range(3):
        k = i + 1
        th = x[k, 4]
        c, s = math.cos(th), math.sin(th)
        vx, vy = x[k, 2], x[k, 3]
        dvx, dvy, dth = m.dx_du[5 * k + 2], m.dx_du[5 * k + 3], m.dx_du[5 * k + 4]
        r_vbx = c * dvx + s * dvy + (-s * vx + c * vy) * dth
        r_vby = -s * dvx + c * dvy + (-c * vx - s * vy) * dth
        rows += [r_vbx, r_vby]
        for obs, dh in [(o, dh_circle) for o in circles] + [(o, dh_ellipse) for o in ellipses]:
            a1, a2 = dh(obs, x[k, 0], x[k, 1])
            b1, b2 = dh(obs, x[i, 0], x[i, 1])
            rows.append(a1 * m.dx_du[5 * k] + a2 * m.dx_du[5 * k + 1]
                        + (form.gamma - 1.0) * (b1 * m.dx_du[5 * i] + b2 * m.dx_du[5 * i + 1]))
        rows.append(2.0 * (x[i, 0] - p[i, 0]) * (m.dx_du[5 * i] - m.dP_du[3 * i])
                    + 2.0 * (x[i, 1] - p[i, 1]) * (m.dx_du[5 * i + 1] - m.dP_du[3 * i + 1]))
        rows.append(m.dP_du[3 * i + 2].copy())
        if form.has_fen:
            sg = 0.0 if p[i, 2] == 0 else form.s_turn * math.copysign(1.0, p[i, 2])
            rows.append(sg * m.dP_du[3 * i + 2] + r_vbx)
    return np.array(rows, dtype=np.float64)


# --------------------------------------------------------------------------------------------------------
# DD callbacks (u in R^6 = (v0, w0, v1, w1, v2, w2), state [x, y, theta])   MPC_DD_sig_step.py:351-572
# --------------------------------------------------------------------------------------------------------
def dd_rollout(xk, u):
    x = np.zeros((4, 3))
    x[0] = np.asarray(xk, dtype=np.float64).ravel()
    u = np.asarray(u, dtype=np.float64).ravel()
    for i in range(3):
        v, w = u[2 * i], u[2 * i + 1]
        x[i + 1, 0] = x[i, 0] + DT * math.cos(x[i, 2]) * v
        x[i + 1, 1] = x[i, 1] + DT * math.sin(x[i, 2]) * v
        x[i + 1, 2] = x[i, 2] + w
    return x


def dd_dx_du(x, u):
    """12 x 6 sensitivity of the stacked states (MPC_DD_sig_step.py:534-566)."""
    d = np.zeros((12, 6))
    for k in range(1, 4):            # state index
        for j in range(k):           # control index
            # d pos_k / d v_j
            d[3 * k + 0, 2 * j] = DT * math.cos(x[j, 2])
            d[3 * k + 1, 2 * j] = DT * math.sin(x[j, 2])
            # d pos_k / d w_j : heading theta_l for l>j shifts with w_j
            sx = sy = 0.0
            for l in range(j + 1, k):
                sx += -u[2 * l] * DT * math.sin(x[l, 2])
                sy += u[2 * l] * DT * math.cos(x[l, 2])
            d[3 * k + 0, 2 * j + 1] = sx
            d[3 * k + 1, 2 * j + 1] = sy
            d[3 * k + 2, 2 * j + 1] = 1.0
    return d


def dd_objective(form: Formulation, xk, goal, last_u, u):
    x = dd_rollout(xk, u)
    u = np.asarray(u, dtype=np.float64).ravel()
    g = np.asarray(goal, dtype=np.float64).ravel()
    prev = np.asarray(last_u, dtype=np.float64).ravel()
    cost = 0.0
    for i in range(3):
        k = i + 1
        d = x[k, 0:2] - g
        tar = math.atan2(g[1] - x[k, 1], g[0] - x[k, 0])
        du = u[2 * i:2 * i + 2] - prev
        cost += form.q * (d @ d) + form.r * (x[k, 2] - tar) ** 2 + form.t_smooth * (du @ du)
        prev = u[2 * i:2 * i + 2]
    d1 = x[1, 0:2] - g
    cost += form.p * (d1 @ d1)
    return float(cost)


def dd_gradient(form: Formulation, xk, goal, last_u, u):
    u = np.asarray(u, dtype=np.float64).ravel()
    x = dd_rollout(xk, u)
    dx = dd_dx_du(x, u)
    g = np.asarray(goal, dtype=np.float64).ravel()
    out = np.zeros(6)
    for k in range(1, 4):
        w = form.q + (form.p if k == 1 else 0.0)
        out += 2.0 * w * ((x[k, 0] - g[0]) * dx[3 * k] + (x[k, 1] - g[1]) * dx[3 * k + 1])
        dx_, dy_ = g[0] - x[k, 0], g[1] - x[k, 1]
        tar = math.atan2(dy_, dx_)
        dtar = (dx_ * (-dx[3 * k + 1]) - dy_ * (-dx[3 * k])) / (dx_ * dx_ + dy_ * dy_)
        out += 2.0 * form.r * (x[k, 2] - tar) * (dx[3 * k + 2] - dtar)
    prev = np.asarray(last_u, dtype=np.float64).ravel()
    eye = np.eye(6)
    for i in range(3):
        for c in range(2):
            cur = u[2 * i + c]
            if i == 0:
                out += 2.0 * form.t_smooth * (cur - prev[c]) * eye[c]
            else:
                out += 2.0 * form.t_smooth * (cur - u[2 * (i - 1) + c]) * (eye[2 * i + c] - eye[2 * (i - 1) + c])
    return out


def dd_bounds(form: Formulation, n_cir: int, n_elp: int):
    """(lb, ub, cl, cu)  MPC_DD_sig_step.py:127-141 (v limits reuse bvx_min/bvx_max = 0.4/0.8)."""
    lb = np.tile([form.bvx_min, -form.ang_max], 3)
    ub = np.tile([form.bvx_max, form.ang_max], 3)
    cl = np.tile(np.r_[np.zeros(n_cir + n_elp), form.bvx_min], 3)
    cu = np.tile(np.r_[np.full(n_cir + n_elp, np.inf), form.bvx_max], 3)
    return lb, ub, cl, cu


def dd_constraints(form: Formulation, xk, circles, ellipses, u):
    u = np.asarray(u, dtype=np.float64).ravel()
    x = dd_rollout(xk, u)
    rows = []
    for i in range(3):
        for cir in circles:
            rows.append(h_circle(cir, x[i + 1, 0], x[i + 1, 1]) + (form.gamma - 1.0) * h_circle(cir, x[i, 0], x[i, 1]))
        for elp in ellipses:
            rows.append(h_ellipse(elp, x[i + 1, 0], x[i + 1, 1]) + (form.gamma - 1.0) * h_ellipse(elp, x[i, 0], x[i, 1]))
        rows.append(form.s_turn * abs(u[2 * i + 1]) + u[2 * i])
    return np.array(rows, dtype=np.float64)


def dd_jacobian(form: Formulation, xk, circles, ellipses, u):
    u = np.asarray(u, dtype=np.float64).ravel()
    x = dd_rollout(xk, u)
    dx = dd_dx_du(x, u)
    rows = []
    for i in range(3):
        k = i + 1
        for obs, dh in [(o, dh_circle) for o in circles] + [(o, dh_ellipse) for o in ellipses]:
            a1, a2 = dh(obs, x[k, 0], x[k, 1])
            b1, b2 = dh(obs, x[i, 0], x[i, 1])
            rows.append(a1 * dx[3 * k] + a2 * dx[3 * k + 1] + (form.gamma - 1.0) * (b1 * dx[3 * i] + b2 * dx[3 * i + 1]))
        r = np.zeros(6)
        r[2 * i] = 1.0
        w = u[2 * i + 1]
        r[2 * i + 1] = 0.0 if w == 0 else form.s_turn * math.copysign(1.0, w)
        rows.append(r)
    return np.array(rows, dtype=np.float64)


# --------------------------------------------------------------------------------------------------------
# planner-level logic around the solve
# --------------------------------------------------------------------------------------------------------
def goal_shift(xk, goal, circles):
    """Detour heuristic (MPC_LIP_sig_step.py:229-253): first circle that is nearer than the goal, within
    3 radii and within 15 deg of the goal bearing rotates the goal by 15 deg away from it."""
    px, py = float(xk[0]), float(xk[1])
    gx, gy = float(goal[0]), float(goal[1])
    d_goal = (px - gx) ** 2 + (py - gy) ** 2
    for cir in circles:
        d_c = (px - cir[0]) ** 2 + (py - cir[1]) ** 2
        if d_c < d_goal and d_c < 9.0 * cir[2] ** 2:
            th = math.atan2(gy - py, gx - px)
            al = math.atan2(cir[1] - py, cir[0] - px)
            d = th - al
            if d < 0 and abs(d) > math.pi:
                d += 2.0 * math.pi
            elif d > 0 and abs(d) > math.pi:
                d -= 2.0 * math.pi
            if abs(d) < math.pi / 12:
                new = th - math.pi / 12 if d < 0 else th + math.pi / 12
                rad = math.sqrt(d_goal)
                return np.array([px + rad * math.cos(new), py + rad * math.sin(new)])
    return np.array([gx, gy])


def select_obs(xk, circles, ellipses, detect_sq: float = 16.0):
    """MPC_LIP_modi.py:325-338 -> boolean keep masks."""
    px, py = float(xk[0]), float(xk[1])
    keep_c = [((px - c[0]) ** 2 + (py - c[1]) ** 2 - c[2] ** 2) <= detect_sq for c in circles]
    keep_e = [((px - e[0]) ** 2 + (py - e[1]) ** 2 - max(e[2], e[3]) ** 2) <= detect_sq for e in ellipses]
    return np.array(keep_c, dtype=bool), np.array(keep_e, dtype=bool)


def sig_step_warm_start(xk, init_guess):
    """MPC_LIP_sig_step.py:185-189."""
    xk = np.asarray(xk, dtype=np.float64).ravel()
    if init_guess is None:
        return np.concatenate([xk, xk, xk])
    g = [np.asarray(v, dtype=np.float64).ravel() for v in init_guess]
    return np.concatenate([g[1], g[2], g[2]])


def u_from_p(xk, z):
    """Representative u for a foot/turn plan z=(p0,p1,p2): u_k := x_{k+1}  (W B = I)."""
    m = model()
    x = np.asarray(xk, dtype=np.float64).ravel().copy()
    z = np.asarray(z, dtype=np.float64).reshape(3, 3)
    u = []
    for i in range(3):
        x = m.A @ x + m.B @ z[i]
        u.append(x.copy())
    return np.concatenate(u)


def p_map():
    """(U, ) with u = U z + u_c(xk): 15 x 9 constant matrix of the representative map above."""
    m = model()
    U = np.zeros((15, 9))
    for j in range(9):
        e = np.zeros(9)
        e[j] = 1.0
        U[:, j] = u_from_p(np.zeros(5), e)
    return U
