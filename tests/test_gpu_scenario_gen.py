"""dcbf_gen_fields / dcbf_gen_states on the GPU against the numpy mirror (oracle/scenario_gen.py): obstacle fields bit for
bit, start states to 2e-15, and the generated batch solves like a host-built one."""
import numpy as np
import pytest
import torch

from oracle import scenario_gen as sg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solvers():
    from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver
    return {f: DcbfSolver(f) for f in ("sig_step", "modi", "dd")}


@pytest.mark.parametrize("form,num,mix,half_gap", [("sig_step", 6, False, 0.8), ("modi", 10, True, 0.4), ("dd", 7, True, 0.4)])
def test_fields_bit_exact(solvers, form, num, mix, half_gap):
    s = solvers[form]
    F, seed = 96, 0xC0FFEE12345
    g = s.gen_fields(F, seed, num, mix=mix, half_gap=half_gap)
    mc, me, md = sg.make_fields(seed, F, num, mix, half_gap=half_gap)
    assert np.array_equal(g["draws"].cpu().numpy(), md)
    assert np.array_equal(g["cir"].cpu().numpy(), mc)
    assert np.array_equal(g["elp"].cpu().numpy(), me)


def test_unbuildable_field_reports(solvers):
    s = solvers["sig_step"]
    g = s.gen_fields(64, 3, 30, half_gap=0.8, install=False)
    assert (g["draws"].cpu().numpy() == -1).all() and torch.isnan(g["cir"]).all()
    assert s.lib.dcbf_gen_fields(s._ctx, 4, 1, 33, 0, 8.5, 1.0, 0.8, 0.4, None, None, None, None) < 0   # more than 32 obstacles


@pytest.mark.parametrize("form", ["sig_step", "modi", "dd"])
def test_states_match_mirror_and_solve(solvers, form):
    from mujoco_lip_mpc_simulation_b200 import scenarios
    s = solvers[form]
    B, F, seed = 2048, 128, 31
    out = scenarios.make_batch_device(s, B, seed, n_fields=F)
    cir, elp = out["cir"].cpu().numpy(), out["elp"].cpu().numpy()
    field = out["field"].cpu().numpy()
    ref = sg.make_states(seed, B, cir, elp, field=field, dd=(form == "dd"), bvy_max=0.3 if form == "sig_step" else 0.35)
    sure = ref["margin"] > 1e-9
    assert sure.mean() > 0.99
    att = out["attempts"].cpu().numpy()
    assert (att >= 1).all() and np.array_equal(att[sure], ref["attempts"][sure])
    assert np.array_equal(out["leg"].cpu().numpy(), ref["leg"])
    np.testing.assert_allclose(out["x0"].cpu().numpy()[sure], ref["x0"][sure], rtol=0, atol=4e-15)
    np.testing.assert_allclose(out["warm"].cpu().numpy()[sure], ref["warm"][sure], rtol=0, atol=4e-15)
    assert np.array_equal(out["goal"].cpu().numpy(), ref["goal"])
    # the batch goes straight into the solver: same answer as the same arrays handed over from the host
    r_dev = s.solve(out["x0"], out["goal"], out["leg"], out["warm"], field=out["field"], last_u=out["last_u"])
    r_host = s.solve(out["x0"].cpu().numpy(), ref["goal"], ref["leg"], out["warm"].cpu().numpy(), field=field,
                     last_u=None if out["last_u"] is None else out["last_u"].cpu().numpy())
    assert torch.equal(r_dev.status, r_host.status) and torch.equal(r_dev.u, r_host.u)
    st = r_dev.status.cpu().numpy()
    assert ((st == 0) | (st == 2)).mean() > 0.99 and (st == 0).mean() > 0.3


def test_million_scenarios_on_device(solvers):
    """config 5's setup without the host: 1 M start states on 65536 fields; launch geometry does not matter (prefix property)."""
    s = solvers["sig_step"]
    big = s.gen_fields(65536, 5, 6, half_gap=0.8)
    st = s.gen_states(1 << 20, 9, field=(torch.arange(1 << 20, device=s.tdev, dtype=torch.int32) % 65536))
    assert int((big["draws"] < 0).sum()) == 0 and int((st["attempts"] < 0).sum()) == 0
    small = s.gen_fields(100, 5, 6, half_gap=0.8, install=False)
    assert torch.equal(small["cir"], big["cir"][:100])
