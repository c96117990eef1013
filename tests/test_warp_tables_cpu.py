"""CPU checks of the pieces the warp kernels take from the host: the lean elementary functions (csrc/dcbf_math.cuh) against
numpy/libm, and the constant tables of csrc/dcbf_warp.cuh (build_warp_tables) against the dense feature-map formulas they
replace.  The tables decide which rows enter which entry of the condensed matrix, so a wrong entry would silently change the
Newton step; here every entry is compared with the dense result."""
import numpy as np

import hostsim_binding as hs


def test_lean_sincos_atan2_match_libm():
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-7, 7, 200000), rng.uniform(-5000, 5000, 50000), [0.0, np.pi / 4, -np.pi / 2, 1e-9]])
    y = rng.uniform(-12, 12, a.size)
    x = rng.uniform(-12, 12, a.size)
    y[::7] *= 1e-6
    x[::11] *= 1e-6
    y[:4], x[:4] = [0.0, 1.0, -1.0, 0.0], [1.0, 0.0, 0.0, -1.0]
    y[4], x[4] = 0.0, 0.0
    sn, cs, at = hs.lean_math(a, y, x)
    assert np.max(np.abs(sn - np.sin(a))) <= 3e-16
    assert np.max(np.abs(cs - np.cos(a))) <= 3e-16
    assert np.max(np.abs(at - np.arctan2(y, x))) <= 5e-16       # 1 ulp at pi
    assert at[4] == 0.0                                            # atan2(0, 0) = 0 like numpy


def test_lean_log_matches_libm():
    """flog (the barrier sums of the warp kernels): < 1 ulp against libm on products of slack gaps, exact at 1"""
    rng = np.random.default_rng(1)
    x = np.concatenate([10.0 ** rng.uniform(-40, 6, 300000), rng.uniform(0.5, 2.0, 100000), 1.0 + rng.uniform(-1e-6, 1e-6, 1000),
                        [1.0, 2.0, 0.5, np.sqrt(2.0), np.nextafter(np.sqrt(2.0), 0), 1e-8 ** 4]])
    got, want = hs.lean_log(x), np.log(x)
    ulp = np.spacing(np.abs(want))
    assert np.max(np.abs(got - want) / np.maximum(ulp, 1e-300)) <= 1.0
    assert hs.lean_log([1.0])[0] == 0.0


def _tri(a, b):
    return a * (a + 1) // 2 + b if a >= b else b * (b + 1) // 2 + a


def _step(a):
    return a >> 1 if a < 6 else a - 6


def test_dot_product_descriptors_cover_the_system_once():
    W = hs.warp_tables()
    seen = {}
    for t, d in enumerate(W["desc"]):
        if d < 0:
            assert t >= 72
            continue
        rowp, rowq, cls, out = d & 0xFF, (d >> 8) & 0xFF, (d >> 16) & 0xF, d >> 20
        assert out not in seen
        seen[out] = (rowp, rowq, cls)
    assert len(seen) == 72
    # condensed matrix: K (packed lower triangle at offset 28) = sum_r (sigma g)[a] g[b]; class = first step that can contribute
    for a in range(9):
        for b in range(a + 1):
            assert seen[28 + _tri(a, b)] == (9 + a, b, max(_step(a), _step(b)))
    # J^T vectors q1, q2, q3 at offset 0: sum_r w_v[r] g[a]
    for v in range(3):
        for a in range(9):
            assert seen[9 * v + a] == (19 + v, a, _step(a))
    # the rounds are sorted by class (lock-step lanes of a round then have similar trip counts)
    cls = [(d >> 16) & 0xF for d in W["desc"][:72]]
    assert cls == sorted(cls)


def test_hessian_table_equals_dense_pullback():
    W = hs.warp_tables()
    T = W["T"]
    rng = np.random.default_rng(1)
    for _ in range(20):
        src = rng.normal(size=27)
        dense = np.zeros((9, 9))
        for kn in range(3):
            h = src[8 * kn: 8 * kn + 8]     # xx, xy, yy, xt, yt, tt, vxt, vyt over (x, y, vx, vy, th)
            H = np.zeros((5, 5))
            H[0, 0], H[0, 1], H[1, 1], H[0, 4], H[1, 4], H[4, 4], H[2, 4], H[3, 4] = h
            H = H + H.T - np.diag(np.diag(H))
            Tk = T[5 * kn: 5 * kn + 5]
            dense += Tk.T @ H @ Tk
        for i in range(3):
            lx, ly = T[15 + 2 * i], T[16 + 2 * i]
            dense += src[24 + i] * (np.outer(lx, lx) + np.outer(ly, ly))
        srcp = np.append(src, 0.0)
        for a in range(9):
            for b in range(a + 1):
                e = _tri(a, b)
                val = sum(W["hc"][t, e] * srcp[W["hs"][t, e]] for t in range(6))
                assert abs(val - dense[a, b]) <= 1e-12 * max(1.0, abs(dense[a, b]))


def test_gradient_coefficients_equal_feature_map():
    """d row / d foot_l = cA[l] * (p0, p1) + cB[l] * (q0, q1) must reproduce the chain rule through the feature map."""
    W = hs.warp_tables()
    T, cab = W["T"], W["cab"]
    rng = np.random.default_rng(2)
    for i in range(3):
        p0, p1, q0, q1 = rng.normal(size=4)
        kn = i + 1
        # D-CBF row of step i: features (x, y) of node i+1 and of node i (node 0 is constant)
        g = p0 * T[5 * (kn - 1) + 0] + p1 * T[5 * (kn - 1) + 1]
        if i > 0:
            g = g + q0 * T[5 * (i - 1) + 0] + q1 * T[5 * (i - 1) + 1]
        for l in range(3):
            assert abs(cab[i, l] * p0 + cab[i, 3 + l] * q0 - g[2 * l]) <= 1e-14
            assert abs(cab[i, l] * p1 + cab[i, 3 + l] * q1 - g[2 * l + 1]) <= 1e-14
        # velocity rows of step i: features (vx, vy) of node i+1
        g = p0 * T[5 * (kn - 1) + 2] + p1 * T[5 * (kn - 1) + 3]
        for l in range(3):
            assert abs(cab[3 + i, l] * p0 - g[2 * l]) <= 1e-14 and abs(cab[3 + i, l] * p1 - g[2 * l + 1]) <= 1e-14
            assert cab[3 + i, 3 + l] == 0.0
        # leg row of step i: features (lx_i, ly_i)
        g = q0 * T[15 + 2 * i] + q1 * T[16 + 2 * i]
        for l in range(3):
            assert abs(cab[6 + i, 3 + l] * q0 - g[2 * l]) <= 1e-14 and abs(cab[6 + i, 3 + l] * q1 - g[2 * l + 1]) <= 1e-14
            assert cab[6 + i, l] == 0.0
    assert np.all(cab[9] == 0.0)


def test_turn_row_descriptors_equal_the_dot_products_they_replace():
    """wp::LipL keeps the three turn rows (gradient e_{6+i}) out of the dot products and adds their contribution in closed
    form (WarpTables::desc_lin_lip): replaying the descriptors on random staged weights must give exactly what the generic dot
    products over those three rows give"""
    W, L = hs.warp_tables(), hs.lin_descriptors()
    desc, lin = W["desc"], L["lin_lip"]
    rng = np.random.default_rng(5)
    N, NST, RP = 9, 22, 38
    ST = np.zeros((NST, RP))
    sig, wts = rng.uniform(0.1, 2.0, 3), rng.standard_normal((3, 3))
    for i in range(3):                       # the turn row of step i is staged column 32 + i
        ST[6 + i, 32 + i] = 1.0              # gradient e_{6+i}
        ST[N + 6 + i, 32 + i] = sig[i]       # sigma * gradient
        ST[2 * N, 32 + i] = sig[i]           # the weights sigma, w1, binv, y
        ST[2 * N + 1:2 * N + 4, 32 + i] = wts[:, i]
    n_used = 0
    for t in range(72):
        d = int(desc[t])
        rowP, rowQ = d & 0xff, (d >> 8) & 0xff
        want = float(ST[rowP, 32:35] @ ST[rowQ, 32:35])          # generic dot product restricted to the turn rows
        dl = int(lin[t])
        got = ST[(dl >> 8) & 0xff, dl >> 16] if dl & 1 else 0.0
        assert got == want, (t, rowP, rowQ, got, want)
        n_used += dl & 1
    assert n_used == 3 + 9 and not lin[72:].any()                 # three diagonal entries, three components of three vectors
