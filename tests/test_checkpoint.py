"""the .npz checkpoint wire format (AIQMCrelease3/checkpoint.py:46-70) and the CSV writer
(utils/writers.py:7-46).  A file written here must load with the reference's own reader (np.load(allow_pickle=True)
+ AINetData(**...)); a file holding pickled jax Arrays (emulated: jax is not installed) must load here."""
import csv
import dataclasses
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from common import CASES, Case

import aiqmc_b200
from aiqmc_b200 import checkpoint


def leaves(t, prefix=""):
    if isinstance(t, dict):
        for k in sorted(t):
            yield from leaves(t[k], prefix + "/" + str(k))
    elif isinstance(t, (list, tuple)):
        for i, v in enumerate(t):
            yield from leaves(v, prefix + "/" + str(i))
    else:
        yield prefix, t


def test_roundtrip_and_reference_reader(tmp_path):
    case = Case(**CASES["C_ecp"], nwalkers=16)
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    opt_state = {"count": np.int32(7), "mu": {"w": np.ones((2, 3))}}
    path = checkpoint.create_save_path(str(tmp_path / "ckpt"))
    assert checkpoint.find_last_checkpoint(path) is None
    checkpoint.save(path, 3, data, case.params, opt_state)
    fname = checkpoint.save(path, 12, data, case.params, opt_state)
    assert fname.endswith("qmcjax_ckpt_000012.npz") and checkpoint.find_last_checkpoint(path) == fname
    # the reference's reader, verbatim logic (checkpoint.py:64-70)
    with open(fname, "rb") as f:
        ck = np.load(f, allow_pickle=True)
        assert sorted(ck.files) == ["data", "opt_state", "params", "t"]
        assert ck["t"].tolist() + 1 == 13
        d = ck["data"].item()
        assert sorted(d) == ["atoms", "charges", "positions", "spins"]
        np.testing.assert_array_equal(d["positions"], case.pos)
        p_ref = ck["params"].tolist()
    # our reader
    t, data2, params2, opt2 = checkpoint.restore(fname, batch_size=16)
    assert t == 13 and isinstance(data2, aiqmc_b200.AINetData)
    np.testing.assert_array_equal(data2.positions, case.pos)
    a, b, c = dict(leaves(case.params)), dict(leaves(params2)), dict(leaves(p_ref))
    assert set(a) == set(b) == set(c)
    for k in a:
        np.testing.assert_array_equal(a[k].numpy(), b[k])
        np.testing.assert_array_equal(a[k].numpy(), c[k])
    assert int(opt2["count"]) == 7
    with pytest.raises(ValueError):
        checkpoint.restore(fname, batch_size=17)
    # a truncated newer file is skipped (checkpoint.py:19-24)
    with open(str(tmp_path / "ckpt" / "qmcjax_ckpt_000099.npz"), "wb") as f:
        f.write(b"PK\x03\x04 not a zip")
    assert checkpoint.find_last_checkpoint(path) == fname


def test_reads_pickled_jax_arrays_without_jax(tmp_path):
    """Emulates what np.savez stores for a pytree of jax Arrays: each leaf pickles as
    jax._src.array._reconstruct_array(fun, args, arr_state, aval_state) with (fun, args, arr_state) numpy's own
    reduce triple.  The fake module exists only while the fixture is written."""
    assert "jax" not in sys.modules

    class FakeJaxArray:
        def __init__(self, v):
            self.v = np.asarray(v)

        def __reduce__(self):
            fun, args, state = self.v.__reduce__()
            return (sys.modules["jax._src.array"]._reconstruct_array, (fun, args, state, {"weak_type": False}))

    mods = {n: types.ModuleType(n) for n in ("jax", "jax._src", "jax._src.array")}
    def _reconstruct_array(*a):
        raise AssertionError("must not be called: jax is not available when reading")
    _reconstruct_array.__module__ = "jax._src.array"
    _reconstruct_array.__qualname__ = "_reconstruct_array"
    mods["jax._src.array"]._reconstruct_array = _reconstruct_array
    sys.modules.update(mods)
    try:
        rng = np.random.default_rng(0)
        params = {"layers": {"streams": [{"single": {"w": FakeJaxArray(rng.normal(size=(5, 4))), "b": FakeJaxArray(rng.normal(size=4))}}]},
                  "envelope": [{"pi": FakeJaxArray(np.ones((1, 3)))}]}
        data = {"positions": FakeJaxArray(rng.normal(size=(1, 8, 12))), "spins": FakeJaxArray(np.ones((1, 8, 4))),
                "atoms": FakeJaxArray(np.zeros((1, 8, 1, 3))), "charges": FakeJaxArray(np.full((1, 8, 1), 4.0))}
        fname = str(tmp_path / "qmcjax_ckpt_000005.npz")
        with open(fname, "wb") as f:
            np.savez(f, t=5, data=data, params=params, opt_state=None)
    finally:
        for n in mods:
            sys.modules.pop(n, None)
    t, d, p, o = checkpoint.restore(fname)
    assert t == 6 and o is None
    assert isinstance(d.positions, np.ndarray) and d.positions.shape == (1, 8, 12)
    np.testing.assert_array_equal(p["layers"]["streams"][0]["single"]["w"], params["layers"]["streams"][0]["single"]["w"].v)
    np.testing.assert_array_equal(p["envelope"][0]["pi"], np.ones((1, 3)))


def test_csv_writer_matches_reference_layout(tmp_path):
    with checkpoint.Writer("train_stats", ["energy", "variance", "pmove"], directory=str(tmp_path / "logs")) as w:
        w.write(0, energy=-5.4, variance=0.3, pmove=0.9)
        w.write(1, energy=-5.41, variance=0.29)
        with pytest.raises(ValueError):
            w.write(2, loss=1.0)
    rows = list(csv.reader(open(tmp_path / "logs" / "train_stats.csv")))
    assert rows == [["t", "energy", "variance", "pmove"], ["0", "-5.4", "0.3", "0.9"], ["1", "-5.41", "0.29", ""]]
