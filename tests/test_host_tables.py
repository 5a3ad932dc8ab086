"""Host-side pieces the GPU tests used to take from the oracle: the product's own spin / pair index tables (A31,
spin_indices.py:5-19,38-46), SystemSpec.from_spins, the workload builders, the Philox restatement's known answers, and
the namedtuple checkpoint round trip."""
import collections
import os

import numpy as np
import pytest

from common import O
from oracle import philox as PH

import aiqmc_b200
from aiqmc_b200 import workloads as W


@pytest.mark.parametrize("spins", [[1., -1., 1., -1.], [1., 1., 1., -1., -1.], [1.] * 5 + [-1.] * 5, [-1., 1.], [1., -1., -1., 1., 1., -1., 1.]])
def test_index_tables_equal_the_oracle_restatement(spins):
    n = len(spins)
    par, anti, npar, nanti = aiqmc_b200.jastrow_indices_ee(spins, n)
    opar, oanti, onpar, onanti = O.jastrow_indices_ee(np.asarray(spins), n)
    assert np.array_equal(par, opar) and np.array_equal(anti, oanti) and (npar, nanti) == (onpar, onanti)
    assert npar + nanti == n * (n - 1) // 2
    up, dn = aiqmc_b200.spin_indices_h(spins)
    oup, odn = O.spin_indices_h(np.asarray(spins))
    assert np.array_equal(up, oup) and np.array_equal(dn, odn)
    spec = aiqmc_b200.SystemSpec.from_spins(np.zeros((1, 3)), [float(n)], spins)
    c = spec.c_struct()
    assert c.n_elec == n and c.n_up == len(up) and c.n_dn == len(dn) and c.n_up_rows == len(up)
    assert [c.sigma[k] for k in range(n)] == list(up) + list(dn)


def test_workloads_describe_the_baseline_configurations():
    sizes = {"c_ae": (6, 1), "c_ecp": (4, 1), "n2": (10, 2), "dmc": (4, 1), "c6h6": (30, 12)}
    for name, (n, a) in sizes.items():
        wl = W.build(name, 16)
        assert (wl.n, wl.a) == (n, a) and wl.pos.shape == (16, 3 * n)
        assert (wl.ecp is not None) == W.SYSTEMS[name]["ecp"]
        lay = aiqmc_b200.lib  # noqa: F841  (the packed-parameter layout needs the library; checked on the GPU box)
    # SURVEY 8(d) numbers
    assert abs(W.flops_walker_step("c_ecp") - 229 * 2410.6667) < 1.0
    assert abs(W.flops_walker_step("c_ae") - 41 * 5496.0) < 1.0
    assert abs(W.flops_walker_step("dmc") - 647 * 2410.6667) < 1.0
    # every rank gets the same parameters and different walkers
    a, b = W.build("n2", 8, rank=0), W.build("n2", 8, rank=1)
    assert np.array_equal(a.params["orbitals"][0]["w"], b.params["orbitals"][0]["w"]) and not np.array_equal(a.pos, b.pos)
    # benzene: 4 electrons start on each carbon, 1 on each hydrogen
    wl = W.build("c6h6", 2000)
    d = np.linalg.norm(wl.pos.reshape(2000, 30, 1, 3) - wl.spec.atoms[None, None], axis=-1).mean(0).argmin(-1)
    assert sorted(np.bincount(d, minlength=12).tolist()) == [1] * 6 + [4] * 6


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    f = lambda c, k: [int(x) for x in PH.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32))]
    assert f([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert f([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert f([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    g1, g2c, u = PH.rng_sweep(5, 1, 10, 50_000, 4, 0.05)
    assert abs(g1.std() / np.sqrt(0.05) - 1) < 0.01 and abs(u.mean() - 0.5) < 0.005 and 0.0 <= u.min() and u.max() < 1.0
    assert np.array_equal(PH.expand_gauss2(g2c)[:, 2, 6:9], g2c[:, 2])


ScaleByAdamState = collections.namedtuple("ScaleByAdamState", ["count", "mu", "nu"])     # optax-style state node


def test_checkpoint_roundtrip_with_namedtuple_opt_state(tmp_path):
    opt = (ScaleByAdamState(count=np.int32(3), mu={"w": np.ones((2, 2))}, nu={"w": np.zeros((2, 2))}), ())
    data = aiqmc_b200.AINetData(positions=np.zeros((2, 6)), spins=np.ones((2, 2)), atoms=np.zeros((2, 1, 3)), charges=np.ones((2, 1)))
    path = aiqmc_b200.checkpoint.save(str(tmp_path), 4, data, {"p": np.arange(3.0)}, opt)
    assert os.path.exists(path)
    t, d2, params, opt2 = aiqmc_b200.checkpoint.restore(path)
    assert t == 5 and np.array_equal(params["p"], np.arange(3.0))
    assert type(opt2[0]).__name__ == "ScaleByAdamState" and int(opt2[0].count) == 3 and np.array_equal(opt2[0].mu["w"], np.ones((2, 2)))
