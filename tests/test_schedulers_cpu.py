"""UniPCMultistepScheduler (editanything_b200.schedulers): diffusers is absent here, so the restatement is pinned by
the method's invariants rather than by the library (parity unpinned against diffusers 0.17.1 itself):
  * solver_order 1 without corrector IS the DDIM eta = 0 update (cldm/ddim_hacked.py:203-231 arithmetic);
  * a model whose x0 prediction is constant makes every order and the corrector agree (all D1 differences vanish);
  * order 2 + corrector converges faster than order 1 on a smooth synthetic denoiser;
  * coefficient_rows() - what the fused device update consumes - reproduces step() exactly."""
import math

import numpy as np
import torch

from editanything_b200.pipeline import DDIMScheduler
from editanything_b200.schedulers import UniPCMultistepScheduler


def _sd_unipc(**kw):
    return UniPCMultistepScheduler.from_config(DDIMScheduler().config, **kw)


def test_from_config_takes_the_sd_schedule_and_timestep_table():
    s = _sd_unipc()
    assert s.config.beta_schedule == "scaled_linear" and abs(s.config.beta_start - 0.00085) < 1e-12
    ref = DDIMScheduler()
    assert np.allclose(s.alphas_cumprod, ref.alphas_cumprod)
    s.set_timesteps(30)
    ts = s.timesteps.tolist()
    assert len(ts) == 30 and ts[0] == 999 and ts[-1] == 33 and all(a > b for a, b in zip(ts, ts[1:]))
    s.set_timesteps(1000)                       # duplicates after rounding are dropped
    assert len(set(s.timesteps.tolist())) == len(s.timesteps)


def test_order1_without_corrector_is_ddim():
    s = _sd_unipc(solver_order=1, disable_corrector=list(range(1000)))
    s.set_timesteps(20)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    acp = s.alphas_cumprod
    ts = s.timesteps.tolist()
    for i, t in enumerate(ts):
        eps = torch.randn(x.shape, generator=g, dtype=torch.float64)
        prev_t = ts[i + 1] if i + 1 < len(ts) else 0
        a, ap = acp[t], acp[prev_t]
        x0 = (x - math.sqrt(1 - a) * eps) / math.sqrt(a)
        ddim = math.sqrt(ap) * x0 + math.sqrt(1 - ap) * eps
        x = s.step(eps, t, x).prev_sample
        assert torch.allclose(x, ddim, atol=1e-10), i


def test_constant_x0_prediction_makes_all_orders_agree():
    g = torch.Generator().manual_seed(1)
    x_init = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
    target = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
    outs = []
    for kw in (dict(solver_order=1, disable_corrector=list(range(1000))), dict(solver_order=2), dict(solver_order=3)):
        s = _sd_unipc(**kw)
        s.set_timesteps(12)
        x = x_init.clone()
        for t in s.timesteps.tolist():
            eps = (x - float(s.alpha_t[t]) * target) / float(s.sigma_t[t])     # the eps that predicts x0 = target
            x = s.step(eps, t, x).prev_sample
        outs.append(x)
    assert torch.allclose(outs[0], outs[1], atol=1e-9) and torch.allclose(outs[0], outs[2], atol=1e-9)


def _solve(order, n, corrector=True):
    """The data-prediction ODE with a smooth x0-predictor x0(x, t) = tanh(0.7 x) * (0.5 + t / 2000), integrated from
    t = 900 to t = 300 in n equal timestep intervals (an interior stretch: the endgame of a real table, a huge
    lambda jump to t = 0 taken at order 1 by lower_order_final, would hide the order of the method)."""
    s = _sd_unipc(solver_order=order, lower_order_final=False, disable_corrector=[] if corrector else list(range(10000)))
    s.set_timesteps(10)
    d = 600 // n
    s.timesteps = torch.tensor([900 - i * d for i in range(n + 1)])
    s._reset()
    x = torch.linspace(-2, 2, 64, dtype=torch.float64).reshape(1, 1, 8, 8)
    for t in s.timesteps.tolist()[:-1]:
        x0 = torch.tanh(0.7 * x) * (0.5 + t / 2000.0)
        eps = (x - float(s.alpha_t[t]) * x0) / float(s.sigma_t[t])
        x = s.step(eps, t, x).prev_sample
    return x


def test_convergence_orders():
    """Halving the step: order-1 predictor error / 2, order-2 predictor / 4, order 2 with the corrector / 8
    (UniPC-p has order p, the corrector raises it to p + 1) - measured 2.0 / 4.0 / 7.9-8.1 on this problem."""
    ref = _solve(2, 600)
    err = {(o, c, n): (_solve(o, n, c) - ref).abs().max().item()
           for o, c in ((1, False), (2, False), (2, True)) for n in (12, 24, 50)}
    for (o, c), lo, hi in (((1, False), 1.8, 2.3), ((2, False), 3.4, 4.8), ((2, True), 6.5, 10.5)):
        r1 = err[(o, c, 12)] / err[(o, c, 24)]
        assert lo < r1 < hi, (o, c, r1)
    assert err[(2, True, 50)] < 0.05 * err[(2, False, 50)] < 0.05 * 0.05 * err[(1, False, 50)] * 20


def test_coefficient_rows_reproduce_step():
    for steps in (2, 5, 30):
        s = _sd_unipc()
        s.set_timesteps(steps)
        rows = s.coefficient_rows()
        g = torch.Generator().manual_seed(steps)
        x = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
        xf = x.clone()
        m1 = m2 = last = torch.zeros_like(x)
        for r, t in zip(rows, s.timesteps.tolist()):
            eps = torch.randn(x.shape, generator=g, dtype=torch.float64)
            x = s.step(eps, t, x).prev_sample
            x0 = (xf - r["sigma"] * eps) / r["alpha"]
            xc = r["kx"] * xf + r["kl"] * last + r["k1"] * m1 + r["k2"] * m2 + r["k0"] * x0
            xf = r["px"] * xc + r["p0"] * x0 + r["p1"] * m1
            m2, m1, last = m1, x0, xc
            assert torch.allclose(x, xf, atol=1e-9), (steps, t)
