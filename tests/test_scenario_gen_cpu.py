"""Scenario generator (csrc/dcbf_gen.cuh) without a GPU: the C++ functions the kernels call, compiled
for the host, against the numpy mirror (oracle/scenario_gen.py) and against fields drawn by the reference generator
(rand_obs.py:31-81, tests/golden/rand_obs_fields.npz)."""
import os

import numpy as np
import pytest

import hostsim_binding as hs
from oracle import scenario_gen as sg

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rand_obs_fields.npz")

# Random123 known-answer vectors for philox4x32-10: (counter, key) -> output
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def field_properties(cir, elp, half_gap, safe_dis, radius=1.0, margin=8.5):
    """What random_circle / random_obs guarantee (rand_obs.py:31-72), on un-inflated or inflated obstacles."""
    discs = [(c[0], c[1], c[2] - safe_dis) for c in cir] + [(e[0], e[1], e[2] - safe_dis) for e in elp]
    for x, y, r in discs:
        assert 0.0 <= x <= margin and 0.0 <= y <= margin and 0.35 - 1e-12 <= r <= radius + 1e-12
        for v in (x, y, r):
            assert abs(v * 100 - round(v * 100)) < 1e-9          # two decimals
    keep = [(10.0, 10.0, 0.3), (0.0, 0.0, 1.0)]
    for i, (x, y, r) in enumerate(discs):
        for (ox, oy, orad) in keep + discs[:i]:
            assert (x - ox) ** 2 + (y - oy) ** 2 - (r + orad + 2 * half_gap) ** 2 >= -1e-9
    for e in elp:
        a, b = e[2] - safe_dis, e[3] - safe_dis
        assert a / 2 - 0.006 <= b <= a + 0.006 and 0.0 <= e[4] <= 3.15
        assert abs(e[4] * 100 - round(e[4] * 100)) < 1e-9 and abs(b * 100 - round(b * 100)) < 1e-9


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(ctr, key, want):
    assert tuple(int(v) for v in sg.philox4x32_10(np.array(ctr, dtype=np.uint32), *key)) == want
    assert tuple(int(v) for v in hs.philox(ctr, *key)) == want


def test_uniforms_are_uniform():
    u0, u1 = sg.uniform2(12345, sg.STREAM_FIELD, np.arange(20000), 7)
    u = np.concatenate([u0, u1])
    assert 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.003
    assert abs(np.corrcoef(u0, u1)[0, 1]) < 0.03


@pytest.mark.parametrize("num,mix,half_gap", [(6, False, 0.8), (10, True, 0.4), (7, True, 0.4), (1, False, 0.8)])
def test_fields_match_the_mirror_bit_for_bit(num, mix, half_gap):
    F, seed = 40, 0x1234_5678_9ABC_DEF0
    cir, elp, draws = hs.gen_fields(F, seed, num, mix, half_gap=half_gap)
    mc, me, md = sg.make_fields(seed, F, num, mix, half_gap=half_gap)
    assert np.array_equal(draws, md) and (draws >= num).all()
    assert np.array_equal(cir, mc) and np.array_equal(elp, me)
    for f in range(F):
        field_properties(cir[f], elp[f], half_gap, 0.4)
    # distinct fields, and a different seed gives a different batch
    assert len({cir[f].tobytes() for f in range(F)}) == F
    assert not np.array_equal(hs.gen_fields(F, seed + 1, num, mix, half_gap=half_gap)[0], cir)


def test_reference_fields_have_the_same_properties():
    G = np.load(GOLD)
    for c in G["cir6"]:
        field_properties(c, np.zeros((0, 5)), 0.8, 0.0)
    for c, e in zip(G["mix_cir"], G["mix_elp"]):
        field_properties(c, e, 0.8, 0.0)
    # same distribution: radius and position statistics of 6-circle fields agree within sampling error
    ours = hs.gen_fields(400, 99, 6, False, half_gap=0.8, safe_dis=0.0)[0]
    ref = G["cir6"]
    assert abs(ours[..., 2].mean() - ref[..., 2].mean()) < 0.05
    assert abs(ours[..., :2].mean() - ref[..., :2].mean()) < 0.4
    # the ellipse conversion keeps b / a in [1/2, 1) and spreads phi over [0, pi] like the reference's
    mixed = hs.gen_fields(400, 98, 6, True, half_gap=0.8, safe_dis=0.0)[1]
    assert abs((mixed[..., 3] / mixed[..., 2]).mean() - (G["mix_elp"][..., 3] / G["mix_elp"][..., 2]).mean()) < 0.05
    assert abs(mixed[..., 4].mean() - np.pi / 2) < 0.15


def test_unbuildable_field_terminates_and_says_so():
    # 30 circles with 0.8 m half gaps do not fit in 8.5 m x 8.5 m: the reference loop would spin forever (rand_obs.py:33-52)
    cir, elp, draws = hs.gen_fields(3, 5, 30, False, half_gap=0.8, stall=200, max_restarts=3)
    assert (draws == -1).all() and np.isnan(cir).all()
    mc, me, md = sg.make_fields(5, 3, 30, False, half_gap=0.8, stall=200, max_restarts=3)
    assert (md == -1).all()


@pytest.mark.parametrize("dd", [False, True])
def test_states_match_the_mirror(dd):
    F, B, seed = 16, 256, 777
    cir, elp, _ = hs.gen_fields(F, seed, 10, True, half_gap=0.4)
    field = (np.arange(B) % F).astype(np.int32)
    got = hs.gen_states(B, seed + 1, cir, elp, field=field, dd=dd, bvy_max=0.35)
    ref = sg.make_states(seed + 1, B, cir, elp, field=field, dd=dd, bvy_max=0.35)
    sure = ref["margin"] > 1e-9                      # no candidate within rounding of the acceptance threshold
    assert sure.mean() > 0.99
    assert np.array_equal(got["attempts"][sure], ref["attempts"][sure]) and (got["attempts"] >= 1).all()
    assert np.array_equal(got["leg"], ref["leg"])
    np.testing.assert_allclose(got["x0"][sure], ref["x0"][sure], rtol=0, atol=2e-15)
    np.testing.assert_allclose(got["warm"][sure], ref["warm"][sure], rtol=0, atol=2e-15)
    assert np.array_equal(got["goal"], ref["goal"])
    # the distribution SURVEY.md 8(d) asks for
    x0 = got["x0"]
    pos, th = x0[:, :2], x0[:, -1]
    assert (pos >= 0).all() and (pos < 8.0).all()
    for b in range(B):
        assert sg.clearance(cir[field[b]], elp[field[b]], pos[b, 0], pos[b, 1]) >= 0.05 - 1e-12
    bearing = np.arctan2(10.0 - pos[:, 1], 10.0 - pos[:, 0])
    assert (np.abs(th - bearing) <= 0.3 + 1e-12).all()
    assert set(np.unique(got["leg"])) == {-1, 1}
    if dd:
        assert np.array_equal(got["warm"], np.tile([0.8, 0.0], (B, 3))) and np.array_equal(got["last_u"], np.tile([0.8, 0.0], (B, 1)))
    else:
        vbx = np.cos(th) * x0[:, 2] + np.sin(th) * x0[:, 3]
        vby = -np.sin(th) * x0[:, 2] + np.cos(th) * x0[:, 3]
        assert (vbx >= 0.4 - 1e-12).all() and (vbx <= 0.8 + 1e-12).all()
        assert (np.abs(vby) >= 0.15 - 1e-12).all() and (np.abs(vby) <= 0.35 + 1e-12).all()
        assert (np.sign(vby) == -got["leg"]).all()
