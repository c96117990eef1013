"""Batched host API over the C ABI: thousands to millions of D-CBF MPC re-plans per call.

`DcbfSolver` owns one `dcbf_ctx` (one formulation, one device).  Device entry points take torch CUDA tensors
(FP64 / int32, contiguous) and enqueue on the current torch stream; `solve_host` takes numpy arrays and goes
through `dcbf_solve_host` (H2D copy, kernel, D2H copy, synchronise).  PyTorch is used for device memory and
streams only.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

FORMS = {"sig_step": 0, "modi": 1, "dd": 2}


def default_params(form) -> _lib.DcbfParams:
    P = _lib.DcbfParams()
    rc = _lib.load().dcbf_default_params(FORMS[form] if isinstance(form, str) else int(form), C.byref(P))
    if rc != 0:
        raise ValueError(f"dcbf_default_params({form}) -> {rc}")
    return P


@dataclass
class SolveResult:
    u: "torch.Tensor | np.ndarray"          # [B,15] (dd: [B,6])  reference decision vector, u_k := x_{k+1}
    x_plan: "torch.Tensor | np.ndarray"     # [B,3,5] (dd: [B,3,3])
    p_plan: "torch.Tensor | np.ndarray | None"   # [B,3,3] foot_x, foot_y, dtheta (LIP only)
    status: "torch.Tensor | np.ndarray"     # [B] int32, Ipopt status integers
    iters: "torch.Tensor | np.ndarray"
    obj: "torch.Tensor | np.ndarray"
    viol: "torch.Tensor | np.ndarray"
    close2goal: "torch.Tensor | np.ndarray"


_PTR_CACHE: dict = {}   # id(ndarray) -> (weak reference, data pointer): ndarray.ctypes / __array_interface__ cost 1.2-1.6 us per access, and a
                        # host-buffer call passes fourteen arrays that are usually the same objects call after call


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.data_ptr()
    hit = _PTR_CACHE.get(id(t))
    if hit is not None and hit[0]() is t:
        return hit[1]
    p = t.__array_interface__["data"][0]
    try:
        if len(_PTR_CACHE) > 512:
            _PTR_CACHE.clear()
        _PTR_CACHE[id(t)] = (weakref.ref(t), p)
    except TypeError:   # (an array subclass without weak references)
        pass
    return p


class DcbfSolver:
    def __init__(self, form="sig_step", device: int | None = None, params: _lib.DcbfParams | None = None, **overrides):
        if not torch.cuda.is_available():
            raise RuntimeError("DcbfSolver needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.P = params if params is not None else default_params(form)
        for k, v in overrides.items():
            setattr(self.P, k, v)
        self.form = int(self.P.formulation)
        self.dd = self.form == 2
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tdev = torch.device("cuda", self.device)
        self._ctx = C.c_void_p()
        # the C ABI restores the caller's current device on exit of every entry point (dcbf_kernels.cu: struct Call)
        rc = self.lib.dcbf_create(C.byref(self.P), self.device, C.byref(self._ctx))
        if rc != 0:
            raise RuntimeError(f"dcbf_create failed ({rc})")
        self.F = self.Kc = self.Ke = 0
        self._keep = []

    # ------------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.dcbf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.dcbf_last_error(self._ctx)
            raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _stream(self):
        return torch.cuda.current_stream(self.tdev).cuda_stream

    def _dev(self, a, dtype):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return a.to(device=self.tdev, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=self.tdev).contiguous()

    @property
    def n(self):
        return 6 if self.dd else 9

    @property
    def nx(self):
        return 3 if self.dd else 5

    @property
    def nu(self):
        return 6 if self.dd else 15

    @property
    def m(self):
        return int(self.lib.dcbf_num_rows(self._ctx))

    @property
    def launches(self):
        return int(self.lib.dcbf_launch_count(self._ctx))

    # ------------------------------------------------------------------------------------------------------
    def set_fields(self, cir, elp=None):
        """cir [F,Kc,3] (cx,cy,r) and elp [F,Ke,5] (cx,cy,a,b,phi), already inflated (obs_cbf of the reference)."""
        cir = self._dev(np.zeros((1, 0, 3)) if cir is None else cir, torch.float64)
        if cir.dim() == 2:
            cir = cir[None]
        F = cir.shape[0]
        elp = self._dev(np.zeros((F, 0, 5)) if elp is None else elp, torch.float64)
        if elp.dim() == 2:
            elp = elp[None]
        assert elp.shape[0] == F
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_set_fields(self._ctx, F, cir.shape[1], _ptr(cir) if cir.numel() else None, elp.shape[1],
                                          _ptr(elp) if elp.numel() else None, self._stream())
        self._check(rc, "dcbf_set_fields")
        self.F, self.Kc, self.Ke = F, int(cir.shape[1]), int(elp.shape[1])
        self._keep = [cir, elp]

    def _inputs(self, x0, goal, leg, field, last_u):
        x0 = self._dev(x0, torch.float64).reshape(-1, self.nx)
        B = x0.shape[0]
        goal = self._dev(goal, torch.float64).reshape(-1, 2)
        if goal.shape[0] != B:
            goal = goal.expand(B, 2).contiguous()
        leg = None if leg is None else self._dev(leg, torch.int32).reshape(-1)
        if leg is not None and leg.shape[0] != B:
            leg = leg.expand(B).contiguous()
        field = None if field is None else self._dev(field, torch.int32).reshape(-1)
        if field is not None and field.shape[0] != B:
            raise ValueError(f"field has {field.shape[0]} entries for a batch of {B} scenarios")
        # field VALUES are checked on the device (an index outside [0, F) makes that scenario return status -13); a host-side
        # min/max would synchronise the stream on every call
        last_u = None if last_u is None else self._dev(last_u, torch.float64).reshape(B, 2)
        return B, x0, goal, leg, field, last_u

    def solve(self, x0, goal, leg, warm, field=None, last_u=None) -> SolveResult:
        """warm=None (LIP formulations): the reference's start vector for init_guess = None, [x_k, x_k, x_k]
        (MPC_LIP_sig_step.py:185-187), formed on the device."""
        B, x0, goal, leg, field, last_u = self._inputs(x0, goal, leg, field, last_u)
        warm = self._warm(warm, B)
        kw = dict(device=self.tdev)
        u = torch.empty((B, self.nu), dtype=torch.float64, **kw)
        xp = torch.empty((B, 3, self.nx), dtype=torch.float64, **kw)
        pp = None if self.dd else torch.empty((B, 3, 3), dtype=torch.float64, **kw)
        st = torch.empty(B, dtype=torch.int32, **kw)
        it = torch.empty(B, dtype=torch.int32, **kw)
        obj = torch.empty(B, dtype=torch.float64, **kw)
        viol = torch.empty(B, dtype=torch.float64, **kw)
        cl = torch.empty(B, dtype=torch.uint8, **kw)
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_solve(self._ctx, B, _ptr(x0), _ptr(goal), _ptr(leg), _ptr(field), _ptr(warm), _ptr(last_u),
                                     _ptr(u), _ptr(xp), _ptr(pp), _ptr(st), _ptr(it), _ptr(obj), _ptr(viol), _ptr(cl),
                                     self._stream())
        self._check(rc, "dcbf_solve")
        return SolveResult(u, xp, pp, st, it, obj, viol, cl.bool())

    def _warm(self, warm, B):
        if warm is None:
            if self.dd:
                raise ValueError("the differential-drive formulation has no start rule of its own: pass warm[B, 6]")
            return None
        return self._dev(warm, torch.float64).reshape(B, self.nu)

    def solve_into(self, B, x0, goal, leg, field, warm, last_u, out: SolveResult):
        """Allocation-free variant for benchmarking: all arguments are resident tensors."""
        rc = self.lib.dcbf_solve(self._ctx, B, _ptr(x0), _ptr(goal), _ptr(leg), _ptr(field), _ptr(warm), _ptr(last_u),
                                 _ptr(out.u), _ptr(out.x_plan), _ptr(out.p_plan), _ptr(out.status), _ptr(out.iters),
                                 _ptr(out.obj), _ptr(out.viol), _ptr(out.close2goal), self._stream())
        self._check(rc, "dcbf_solve")

    def evaluate(self, x0, goal, leg, z, lam=None, field=None, last_u=None, want_hess=True):
        B, x0, goal, leg, field, last_u = self._inputs(x0, goal, leg, field, last_u)
        n, m = self.n, self.m
        z = self._dev(z, torch.float64).reshape(B, n)
        lam = None if lam is None else self._dev(lam, torch.float64).reshape(B, m)
        kw = dict(device=self.tdev, dtype=torch.float64)
        f, grad = torch.empty(B, **kw), torch.empty((B, n), **kw)
        c, jac = torch.empty((B, m), **kw), torch.empty((B, m, n), **kw)
        cl, cu = torch.empty((B, m), **kw), torch.empty((B, m), **kw)
        hess = torch.empty((B, n, n), **kw) if want_hess else None
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_eval(self._ctx, B, _ptr(x0), _ptr(goal), _ptr(leg), _ptr(field), _ptr(last_u), _ptr(z),
                                    _ptr(lam), _ptr(f), _ptr(grad), _ptr(c), _ptr(jac), _ptr(cl), _ptr(cu), _ptr(hess),
                                    self._stream())
        self._check(rc, "dcbf_eval")
        return dict(f=f, grad=grad, c=c, jac=jac, cl=cl, cu=cu, hess=hess)

    def setup_info(self, x0, goal, field=None):
        """Obstacle selection and detour goal of every scenario, as the solve kernels apply them (dcbf_setup_info).
        Returns dict(mask[B] int64 bit field: circles 0..Kc-1, ellipses Kc.., count[B], goal_eff[B,2])."""
        B, x0, goal, _, field, _ = self._inputs(x0, goal, None, field, None)
        mask = torch.empty(B, dtype=torch.int32, device=self.tdev)
        count = torch.empty(B, dtype=torch.int32, device=self.tdev)
        ge = torch.empty((B, 2), dtype=torch.float64, device=self.tdev)
        rc = self.lib.dcbf_setup_info(self._ctx, B, _ptr(x0), _ptr(goal), _ptr(field), _ptr(mask), _ptr(count), _ptr(ge), self._stream())
        self._check(rc, "dcbf_setup_info")
        return dict(mask=mask.to(torch.int64) & 0xFFFFFFFF, count=count, goal_eff=ge)

    def rollout(self, steps, x0, goal, leg, field=None, want_traj=True):
        if self.dd:
            raise ValueError("the closed-loop rollout belongs to the LIP formulations")
        B, x0, goal, leg, field, _ = self._inputs(x0, goal, leg, field, None)
        kw = dict(device=self.tdev)
        xf = torch.empty((B, 5), dtype=torch.float64, **kw)
        sd = torch.empty(B, dtype=torch.int32, **kw)
        ni = torch.empty(B, dtype=torch.int32, **kw)
        ti = torch.empty(B, dtype=torch.int32, **kw)
        traj = torch.empty((B, steps, 8), dtype=torch.float64, **kw) if want_traj else None
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_rollout(self._ctx, B, int(steps), _ptr(x0), _ptr(goal), _ptr(leg), _ptr(field), _ptr(xf),
                                       _ptr(sd), _ptr(ni), _ptr(ti), _ptr(traj), self._stream())
        self._check(rc, "dcbf_rollout")
        return dict(x_final=xf, steps_done=sd, n_infeasible=ni, total_iters=ti, traj=traj)

    def tick(self, glo_pos, glo_vel, glo_hd, glo_p, t_rest, goal, leg, prev_plan=None, mode=None, field=None, want_pos_det=True):
        """One control tick per scenario on the device (dcbf_tick): LIP prediction to the end of the running step
        (MPCCBF.get_next_states), warm start from the previous plan (mode 0 verbatim, 1 shifted, 2 / no previous plan:
        [x_next] * 3 -- data_procs/logger_mpc.py:326-333), re-plan, dense plan trajectory pos_det [B,126,2]."""
        assert not self.dd, "the tick path belongs to the LIP formulations"
        pos = self._dev(glo_pos, torch.float64).reshape(-1, 2)
        B = pos.shape[0]
        vel = self._dev(glo_vel, torch.float64).reshape(B, 2)
        hd = self._dev(glo_hd, torch.float64).reshape(B)
        gp = self._dev(glo_p, torch.float64).reshape(B, 3)
        tr = self._dev(t_rest, torch.float64).reshape(-1)
        if tr.shape[0] != B:
            tr = tr.expand(B).contiguous()
        _, _, goal, leg, field, _ = self._inputs(torch.zeros((B, 5), dtype=torch.float64, device=self.tdev), goal, leg, field, None)
        prev = None if prev_plan is None else self._dev(prev_plan, torch.float64).reshape(B, 15)
        md = None if (mode is None or prev is None) else self._dev(mode, torch.uint8).reshape(-1)
        if md is not None and md.shape[0] != B:
            md = md.expand(B).contiguous()
        if prev is not None and md is None:
            md = torch.zeros(B, dtype=torch.uint8, device=self.tdev)
        kw = dict(device=self.tdev)
        xn = torch.empty((B, 5), dtype=torch.float64, **kw)
        warm = torch.empty((B, 15), dtype=torch.float64, **kw)
        u = torch.empty((B, 15), dtype=torch.float64, **kw)
        xp = torch.empty((B, 3, 5), dtype=torch.float64, **kw)
        pp = torch.empty((B, 3, 3), dtype=torch.float64, **kw)
        st = torch.empty(B, dtype=torch.int32, **kw)
        it = torch.empty(B, dtype=torch.int32, **kw)
        obj = torch.empty(B, dtype=torch.float64, **kw)
        viol = torch.empty(B, dtype=torch.float64, **kw)
        cl = torch.empty(B, dtype=torch.uint8, **kw)
        pd = torch.empty((B, 126, 2), dtype=torch.float64, **kw) if want_pos_det else None
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_tick(self._ctx, B, _ptr(pos), _ptr(vel), _ptr(hd), _ptr(gp), _ptr(tr), _ptr(goal), _ptr(leg), _ptr(field),
                                    _ptr(prev), _ptr(md), _ptr(xn), _ptr(warm), _ptr(u), _ptr(xp), _ptr(pp), _ptr(st), _ptr(it),
                                    _ptr(obj), _ptr(viol), _ptr(cl), _ptr(pd), self._stream())
        self._check(rc, "dcbf_tick")
        return dict(x_next=xn, warm=warm, plan=SolveResult(u, xp, pp, st, it, obj, viol, cl.bool()), pos_det=pd)

    def alip_foot(self, x_alip, y_alip, time, support, speed, speed_stride=1, H=1.0, T=0.4, m=45.0, W=0.2):
        """ALIP one-step foot placement on the device (dcbf_alip_foot); `speed` may be the `u` tensor of a DD solve with
        speed_stride=6 (v_0 of every scenario) so that the two calls chain on the stream."""
        xa = self._dev(x_alip, torch.float64).reshape(-1, 2)
        B = xa.shape[0]
        ya = self._dev(y_alip, torch.float64).reshape(B, 2)
        tm = self._dev(time, torch.float64).reshape(-1)
        if tm.shape[0] != B:
            tm = tm.expand(B).contiguous()
        sup = self._dev(support, torch.int32).reshape(-1)
        if sup.shape[0] != B:
            sup = sup.expand(B).contiguous()
        sp = self._dev(speed, torch.float64).reshape(-1)
        assert sp.numel() >= (B - 1) * speed_stride + 1
        kw = dict(device=self.tdev, dtype=torch.float64)
        foot, am, nxt = torch.empty((B, 2), **kw), torch.empty((B, 2), **kw), torch.empty((B, 4), **kw)
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_alip_foot(self._ctx, B, _ptr(xa), _ptr(ya), _ptr(tm), _ptr(sup), _ptr(sp), int(speed_stride),
                                         float(H), float(T), float(m), float(W), _ptr(foot), _ptr(am), _ptr(nxt), self._stream())
        self._check(rc, "dcbf_alip_foot")
        return dict(foot=foot, am=am, next=nxt)

    def veldes_foot(self, x_state=None, leg=None, vel_des=None, vx_max=0.6, step_gap=0.3):
        """MPCCBF.alip_des_vel + MPCCBF.cal_foot_with_veldes for B scenarios on the device (dcbf_veldes_foot): the desired
        end-of-step velocity for walking speed vx_max with stance sign leg (or vel_des as given) and the foothold that reaches it
        from x_state [B,5].  Returns dict(vel_des[B,2], foot[B,2] | None)."""
        assert not self.dd, "the velocity-tracking foothold belongs to the LIP formulations"
        xs = None if x_state is None else self._dev(x_state, torch.float64).reshape(-1, 5)
        lg = None if leg is None else self._dev(leg, torch.int32).reshape(-1)
        vd = None if vel_des is None else self._dev(vel_des, torch.float64).reshape(-1, 2)
        B = next(t.shape[0] for t in (xs, vd, lg) if t is not None)
        if lg is not None and lg.shape[0] != B:
            lg = lg.expand(B).contiguous()
        out_v = torch.empty((B, 2), dtype=torch.float64, device=self.tdev)
        foot = None if xs is None else torch.empty((B, 2), dtype=torch.float64, device=self.tdev)
        rc = self.lib.dcbf_veldes_foot(self._ctx, B, _ptr(xs), _ptr(lg), _ptr(vd), float(vx_max), float(step_gap), _ptr(out_v), _ptr(foot),
                                       self._stream())
        self._check(rc, "dcbf_veldes_foot")
        return dict(vel_des=out_v, foot=foot)

    def heading_input(self, cur_hd, nex_turn, x_plan=None, mpc_hds=None, glo_p=None):
        """Logger.tube_func + Logger.avg_hd for B scenarios (dcbf_heading_input).  nex_turn [B] (device tensor) is updated in
        place.  The plan headings come from x_plan [B,3,5] (a dcbf_tick / solve output, read in place) or mpc_hds [B,3]; the
        result is written into glo_p[:, 2] when glo_p [B,3] is given (the input of the next tick) and returned."""
        cur = self._dev(cur_hd, torch.float64).reshape(-1)
        B = cur.shape[0]
        assert isinstance(nex_turn, torch.Tensor) and nex_turn.is_cuda and nex_turn.dtype == torch.float64 and nex_turn.is_contiguous()
        if x_plan is not None:
            src = self._dev(x_plan, torch.float64).reshape(B, 15)
            base, stride, step = src.data_ptr() + 4 * 8, 15, 5
        else:
            src = self._dev(mpc_hds, torch.float64).reshape(B, 3)
            base, stride, step = src.data_ptr(), 3, 1
        if glo_p is not None:
            assert isinstance(glo_p, torch.Tensor) and glo_p.is_cuda and glo_p.is_contiguous() and glo_p.shape == (B, 3)
            assert glo_p.dtype == torch.float64 and glo_p.device == self.tdev, "glo_p must be an FP64 tensor on the solver's device"
            assert nex_turn.device == self.tdev
            out, optr, ostride = glo_p[:, 2], glo_p.data_ptr() + 2 * 8, 3
        else:
            out = torch.empty((B,), device=self.tdev, dtype=torch.float64)
            optr, ostride = out.data_ptr(), 1
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_heading_input(self._ctx, B, _ptr(cur), _ptr(nex_turn), base, stride, step, optr, ostride, self._stream())
        self._check(rc, "dcbf_heading_input")
        return out

    def gen_fields(self, F, seed, num, mix=False, margin=8.5, radius=1.0, half_gap=0.8, safe_dis=0.4, install=True):
        """F random obstacle fields on the device (dcbf_gen_fields: rand_obs.py:31-81 batched, with bounded restarts).
        Returns dict(cir[F,Kc,3], elp[F,Ke,5], draws[F]); install=True also makes them the context's fields."""
        Kc, Ke = ((num + 1) // 2, num // 2) if mix else (num, 0)
        kw = dict(device=self.tdev, dtype=torch.float64)
        cir, elp = torch.empty((F, Kc, 3), **kw), torch.empty((F, Ke, 5), **kw)
        draws = torch.empty((F,), device=self.tdev, dtype=torch.int32)
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_gen_fields(self._ctx, int(F), int(seed) & 0xFFFFFFFFFFFFFFFF, int(num), int(bool(mix)), float(margin),
                                          float(radius), float(half_gap), float(safe_dis), _ptr(cir), _ptr(elp) if Ke else None,
                                          _ptr(draws), self._stream())
        self._check(rc, "dcbf_gen_fields")
        if install:
            self.set_fields(cir, elp)
        return dict(cir=cir, elp=elp, draws=draws)

    def gen_states(self, B, seed, field=None, goal=(10.0, 10.0), bvy_max=0.0):
        """B start states on the context's fields (dcbf_gen_states).  Returns dict(x0, goal, leg, warm, last_u|None, attempts)."""
        fld = None if field is None else self._dev(field, torch.int32).reshape(B)
        kw = dict(device=self.tdev, dtype=torch.float64)
        x0, g, warm = torch.empty((B, self.nx), **kw), torch.empty((B, 2), **kw), torch.empty((B, self.nu), **kw)
        leg = torch.empty((B,), device=self.tdev, dtype=torch.int32)
        att = torch.empty((B,), device=self.tdev, dtype=torch.int32)
        last_u = torch.empty((B, 2), **kw) if self.dd else None
        with torch.cuda.device(self.tdev):
            rc = self.lib.dcbf_gen_states(self._ctx, int(B), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(fld) if fld is not None else None,
                                          float(goal[0]), float(goal[1]), float(bvy_max), _ptr(x0), _ptr(g), _ptr(leg), _ptr(warm),
                                          _ptr(last_u) if last_u is not None else None, _ptr(att), self._stream())
        self._check(rc, "dcbf_gen_states")
        return dict(x0=x0, goal=g, leg=leg, warm=warm, last_u=last_u, attempts=att, field=fld)

    # ------------------------------------------------------------------------------------------------------
    def set_fields_host(self, cir, elp=None):
        cir = np.ascontiguousarray(np.zeros((1, 0, 3)) if cir is None else cir, dtype=np.float64)
        if cir.ndim == 2:
            cir = cir[None]
        F = cir.shape[0]
        elp = np.ascontiguousarray(np.zeros((F, 0, 5)) if elp is None else elp, dtype=np.float64)
        if elp.ndim == 2:
            elp = elp[None]
        rc = self.lib.dcbf_set_fields_host(self._ctx, F, cir.shape[1], _ptr(cir) if cir.size else None, elp.shape[1],
                                           _ptr(elp) if elp.size else None)
        self._check(rc, "dcbf_set_fields_host")
        self.F, self.Kc, self.Ke = F, int(cir.shape[1]), int(elp.shape[1])

    def solve_host(self, x0, goal, leg, warm, field=None, last_u=None, out: SolveResult | None = None, wait: bool = True) -> SolveResult:
        """numpy in / numpy out through dcbf_solve_host (copies + kernel + copies, synchronous).  wait=False enqueues only
        (dcbf_solve_host_async: every buffer, `out` included, must be page-locked and stay untouched until wait())."""
        def ready(a, dt, n):   # already what the C side takes: a C-contiguous array of the right type and size (no conversion, no copy)
            return isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.size == n
        B = x0.shape[0] if isinstance(x0, np.ndarray) and x0.ndim == 2 else -1
        if warm is None and self.dd:
            raise ValueError("the differential-drive formulation has no start rule of its own: pass warm[B, 6]")
        if not (B > 0 and ready(x0, np.float64, B * self.nx) and ready(goal, np.float64, 2 * B) and (warm is None or ready(warm, np.float64, B * self.nu))
                and (leg is None or ready(leg, np.int32, B)) and (field is None or ready(field, np.int32, B))
                and (last_u is None or ready(last_u, np.float64, 2 * B))):
            f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
            x0 = f64(x0).reshape(-1, self.nx)
            B = x0.shape[0]
            goal = f64(np.broadcast_to(f64(goal).reshape(-1, 2), (B, 2)))
            leg = None if leg is None else np.ascontiguousarray(np.broadcast_to(np.asarray(leg, dtype=np.int32).reshape(-1), (B,)))
            field = None if field is None else np.ascontiguousarray(field, dtype=np.int32).reshape(-1)
            if field is not None and field.shape[0] != B:
                raise ValueError(f"field has {field.shape[0]} entries for a batch of {B} scenarios")
            warm = None if warm is None else f64(warm).reshape(B, self.nu)
            last_u = None if last_u is None else f64(last_u).reshape(B, 2)
        if out is None:
            out = SolveResult(np.empty((B, self.nu)), np.empty((B, 3, self.nx)), None if self.dd else np.empty((B, 3, 3)),
                              np.empty(B, np.int32), np.empty(B, np.int32), np.empty(B), np.empty(B), np.empty(B, np.uint8))
        fn = self.lib.dcbf_solve_host_async if wait is False else self.lib.dcbf_solve_host
        rc = fn(self._ctx, B, _ptr(x0), _ptr(goal), _ptr(leg), _ptr(field), _ptr(warm), _ptr(last_u),
                _ptr(out.u), _ptr(out.x_plan), _ptr(out.p_plan), _ptr(out.status), _ptr(out.iters),
                _ptr(out.obj), _ptr(out.viol), _ptr(out.close2goal))
        self._check(rc, "dcbf_solve_host_async" if wait is False else "dcbf_solve_host")
        if wait is False:
            self._pending = (x0, goal, leg, field, warm, last_u, out)   # keep the buffers alive until wait()
        return out

    def wait(self):
        """dcbf_wait: returns when every solve_host(..., wait=False) of this solver has delivered its results."""
        rc = self.lib.dcbf_wait(self._ctx)
        self._check(rc, "dcbf_wait")
        self._pending = None

    def fp64_peak_tflops(self, repeats: int = 3) -> float:
        return float(self.lib.dcbf_fp64_peak_tflops(self._ctx, repeats))


class HostPipeline:
    """A stream of host-buffer batches through `lanes` contexts used round-robin (dcbf_solve_host_async / dcbf_wait): the drain of
    one batch overlaps the head of the next, so a caller with many small batches gets the throughput of a large one.  Every buffer
    must be page-locked (`pin(...)` below); submit() returns the SolveResult the batch will be written to and the lane it runs
    on; its contents are valid after that lane's wait() (or after drain())."""

    def __init__(self, form="sig_step", lanes: int = 3, device: int | None = None, **overrides):
        self.solvers = [DcbfSolver(form, device=device, **overrides) for _ in range(lanes)]
        self._busy = [False] * lanes
        self._next = 0

    @staticmethod
    def pin(a):
        """page-locked copy of a numpy array (what the copy-free host path needs)"""
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    def set_fields_host(self, cir, elp=None):
        for s in self.solvers:
            s.set_fields_host(cir, elp)

    def submit(self, x0, goal, leg, warm, field=None, last_u=None, out: SolveResult | None = None):
        lane = self._next
        self._next = (lane + 1) % len(self.solvers)
        s = self.solvers[lane]
        if self._busy[lane]:
            s.wait()
        if out is None:
            B = np.asarray(x0).reshape(-1, s.nx).shape[0]
            out = SolveResult(self.pin(np.empty((B, s.nu))), self.pin(np.empty((B, 3, s.nx))), None if s.dd else self.pin(np.empty((B, 3, 3))),
                              self.pin(np.empty(B, np.int32)), self.pin(np.empty(B, np.int32)), self.pin(np.empty(B)), self.pin(np.empty(B)),
                              self.pin(np.empty(B, np.uint8)))
        s.solve_host(x0, goal, leg, warm, field=field, last_u=last_u, out=out, wait=False)
        self._busy[lane] = True
        return out, lane

    def wait(self, lane: int):
        if self._busy[lane]:
            self.solvers[lane].wait()
            self._busy[lane] = False

    def drain(self):
        for lane in range(len(self.solvers)):
            self.wait(lane)
