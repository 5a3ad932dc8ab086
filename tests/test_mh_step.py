"""all-electron Metropolis-Hastings (AIQMCrelease2/MonteCarloSample/mcstep.py:12-124).
CPU: the oracle restatement against closed forms; GPU: aiqmc_mh_step / make_mcmc_step against the oracle with the
same noise and uniforms -- accept masks bit-exact, positions and 2 log|psi| identical, pmove equal."""
import math

import numpy as np
import pytest
import torch

from common import CASES, Case, O

import aiqmc_b200


def test_oracle_harmonic_mean_and_gaussian_logprob_closed_forms():
    atoms = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 2.0]])
    x = torch.tensor([[[[0.0, 0.0, 1.0]], [[3.0, 0.0, 0.0]]]])                  # (1,2,1,3)
    h = O._harmonic_mean(x, atoms)
    assert h.shape == (1, 2, 1, 1)
    np.testing.assert_allclose(h[0, 0, 0, 0], 1.0)                               # 1 / mean(1/1, 1/1)
    np.testing.assert_allclose(h[0, 1, 0, 0], 2.0 / (1 / 3.0 + 1 / math.sqrt(13.0)))
    sig = 0.3 * h
    lp = O._log_prob_gaussian(x + 0.1, x, sig)
    ref = sum(-0.5 * 3 * 0.01 / float(s) ** 2 - 3 * math.log(float(s)) for s in sig.reshape(-1))
    np.testing.assert_allclose(float(lp), ref, rtol=1e-13)


def test_oracle_mh_samples_a_known_density():
    """f = -|x| per electron (hydrogen-like, the function ferminet/tests use): <r> of |psi|^2 = exp(-2r) is 3/2."""
    torch.manual_seed(0)
    B, steps = 2048, 60
    atoms = torch.zeros(1, 3)
    f = lambda params, x, *_: -x.reshape(x.shape[0], -1, 3).norm(dim=-1).sum(-1)
    data = O.AINetData(positions=torch.randn(B, 3, dtype=torch.float64), spins=None, atoms=None, charges=None)
    step = O.make_mcmc_step(f, B, steps=steps, atoms=atoms)
    rand = dict(noise=torch.randn(steps, B, 3, dtype=torch.float64), u=torch.rand(steps, B, dtype=torch.float64))
    data, pmove, masks, lp = step(None, data, rand, 0.6)
    assert 0.3 < float(pmove) < 0.9
    np.testing.assert_allclose(float(data.positions.norm(dim=-1).mean()), 1.5, atol=0.08)
    np.testing.assert_allclose(lp.numpy(), -2.0 * data.positions.norm(dim=-1).numpy(), rtol=1e-13)
    w, pm = O.update_mcmc_width(20, 0.6, 20, 0.9, np.full(20, 0.9))
    assert abs(w - 0.66) < 1e-12
    w, pm = O.update_mcmc_width(20, 0.6, 20, 0.1, np.full(20, 0.1))
    assert abs(w - 0.6 / 1.1) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name,B", [("C_ae", 300), ("N2_ecp", 65), ("h2like", 1000)])
def test_mcmc_step_matches_oracle_bit_exact_accepts(name, B):
    steps, width = 4, 0.25
    case = Case(**CASES[name], nwalkers=B)
    rng = np.random.default_rng(4)
    noise = rng.normal(size=(steps, B, 3 * case.n))
    u = rng.uniform(size=(steps, B))
    net = aiqmc_b200.make_ai_net(**case.kw)
    step = aiqmc_b200.make_mcmc_step(net.apply, B, steps=steps, atoms=case.t_atoms)
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    new_data, pmove = step(case.params, data, dict(noise=torch.tensor(noise).cuda(), u=torch.tensor(u).cuda()), width)
    logabs = lambda p, x, s, a, c: case.net.apply(p, x, case.t_spins, case.t_atoms)[1]
    ostep = O.make_mcmc_step(logabs, B, steps=steps, atoms=case.t_atoms)
    odata, opmove, masks, olp = ostep(case.params, case.oracle_data(), dict(noise=torch.tensor(noise), u=torch.tensor(u)), width)
    assert 0.0 < float(opmove) < 1.0
    assert round(float(pmove) * steps * B) == round(float(opmove) * steps * B)   # the same integer accept count
    np.testing.assert_allclose(float(pmove), float(opmove), rtol=1e-14)
    np.testing.assert_allclose(new_data.positions.cpu().numpy(), odata.positions.numpy(), rtol=1e-12, atol=1e-12)
    # the single-step entry point: accept mask of the first step, bit for bit
    eng = net.apply.bind(case.params, case.t_atoms)
    pos = torch.tensor(case.pos).cuda()
    lp = (2.0 * eng.psi(pos, mode=0)[1]).contiguous()
    cnt = torch.zeros((), dtype=torch.int64, device="cuda")
    acc = eng.mh_step(pos, lp, torch.tensor(noise[0]).cuda(), torch.tensor(u[0]).cuda(), width, cnt, want_accept=True)
    assert np.array_equal(acc.cpu().numpy().astype(bool), masks[0].numpy())
    assert int(cnt) == int(masks[0].sum())
    # 2 log|psi| carried along equals a fresh evaluation at the new positions (north_star: 1e-6 relative)
    np.testing.assert_allclose(lp.cpu().numpy(), 2.0 * eng.psi(pos, mode=0)[1].cpu().numpy(), rtol=1e-12)


@pytest.mark.gpu
def test_mh_empty_batch_and_width_update():
    case = Case(**CASES["C_ae"], nwalkers=4)
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params)
    cnt = torch.zeros((), dtype=torch.int64, device="cuda")
    pos = torch.zeros((0, 18), dtype=torch.float64, device="cuda")
    eng.mh_step(pos, torch.zeros(0, dtype=torch.float64, device="cuda"), pos, torch.zeros(0, device="cuda"), 0.1, cnt)
    assert int(cnt) == 0
    w, pm = aiqmc_b200.update_mcmc_width(20, 0.6, 20, torch.tensor(0.9), np.full(20, 0.9))
    assert abs(w - 0.66) < 1e-12
