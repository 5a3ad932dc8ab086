"""correlated-sampling estimators (correlatedsamples/corrsamples.py:23-47, jacobianWeights.py:22-51)
and the ccECP file reader that pseudopotential/readpp.py left unfinished."""
import numpy as np
import pytest
import torch

from common import C_ECP, O

import aiqmc_b200

# pseudopotential/C.ccECP.nwchem:1-7 verbatim (published ccECP for carbon)
C_CCECP = """C nelec 2
C ul
1 14.43502 4.00000
3 8.39889 57.74008
2 7.38188 -25.81955
C S
2 7.76079 52.13345
"""


def test_ecp_reader_reproduces_the_tables_the_reference_hard_codes():
    """Known answer: example/single_atom_C/single_atom_C.py:13-23 (the same numbers tests/common.C_ECP carries)."""
    t = aiqmc_b200.read_ecp_nwchem(C_CCECP, ["C"])
    np.testing.assert_array_equal(t["rn_local"], [[1.0, 3.0, 2.0]])
    np.testing.assert_array_equal(t["local_coes"], [[4.00000, 57.74008, -25.81955]])
    np.testing.assert_array_equal(t["local_exps"], [[14.43502, 8.39889, 7.38188]])
    np.testing.assert_array_equal(t["rn_non_local"], [[[2.0, 2.0], [2.0, 2.0], [2.0, 2.0]]])
    np.testing.assert_array_equal(t["non_local_coes"], [[[52.13345, 0], [0, 0], [0, 0]]])
    np.testing.assert_array_equal(t["non_local_exps"], [[[7.76079, 0], [0, 0], [0, 0]]])
    assert t["list_l"] == 2 and t["nelec_core"].tolist() == [2]
    for k, v in C_ECP.items():
        np.testing.assert_array_equal(t[k], v)
    two = aiqmc_b200.read_ecp_nwchem(C_CCECP, ["C", "C"])                       # readpp.py:6-7: symbol = ['C', 'C']
    assert two["rn_local"].shape == (2, 3) and np.array_equal(two["local_coes"][0], two["local_coes"][1])
    ecp = aiqmc_b200.make_ecp(2, list_l=two["list_l"], **{k: two[k] for k in C_ECP})
    assert ecp.k_loc == 3 and ecp.n_l == 3
    with pytest.raises(ValueError):
        aiqmc_b200.read_ecp_nwchem(C_CCECP, ["N"])
    with pytest.raises(ValueError):
        aiqmc_b200.read_ecp_nwchem(C_CCECP + "C P\n2 1.0 2.0\n2 1.1 2.1\n2 1.2 2.2\n", ["C"])


def test_oracle_space_warp_limits():
    """Closed forms: a rigid translation of all nuclei translates every electron by the same vector; an electron
    sitting (almost) on a nucleus follows that nucleus."""
    atoms = torch.tensor([[0.0, 0.0, 0.0], [2 / 3, 1 / 3, 0.0]])
    shift = torch.tensor([0.2, -0.1, 0.05])
    pos = torch.tensor(np.random.default_rng(0).normal(size=12))
    np.testing.assert_allclose(O.correlated_samples(atoms, atoms + shift, pos).reshape(4, 3).numpy(),
                               pos.reshape(4, 3).numpy() + shift.numpy(), rtol=1e-13)
    new_atoms = atoms.clone()
    new_atoms[1] += torch.tensor([0.1, 0.0, 0.0])
    pos2 = torch.cat([atoms[1] + 1e-6, torch.tensor([50.0, 0.0, 0.0])])
    out = O.correlated_samples(atoms, new_atoms, pos2).reshape(2, 3)
    np.testing.assert_allclose(out[0].numpy(), (atoms[1] + 1e-6 + torch.tensor([0.1, 0, 0])).numpy(), atol=1e-12)
    j = O.weights_jacobian(pos, atoms, atoms)
    assert float(j) == 1.0                                                       # no displacement: unit weight


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,a", [(1, 8, 2), (257, 4, 1), (4096, 10, 3)])
def test_correlated_samples_and_jacobian_match_oracle(B, n, a):
    rng = np.random.default_rng(B)
    atoms = rng.normal(size=(a, 3))
    new_atoms = atoms + 0.1 * rng.normal(size=(a, 3))
    pos = rng.normal(size=(B, 3 * n)) * 1.5
    out = aiqmc_b200.correlated_samples(atoms, new_atoms, torch.tensor(pos)).cpu().numpy()
    jac = aiqmc_b200.weights_jacobian(torch.tensor(pos), atoms, new_atoms).cpu().numpy()
    for b in range(min(B, 40)):
        ref = O.correlated_samples(torch.tensor(atoms), torch.tensor(new_atoms), torch.tensor(pos[b]))
        np.testing.assert_allclose(out[b], ref.numpy(), rtol=1e-12, atol=1e-13)
        jr = O.weights_jacobian(torch.tensor(pos[b]), torch.tensor(atoms), torch.tensor(new_atoms))
        np.testing.assert_allclose(jac[b], float(jr), rtol=1e-9)
    # size-independent property at full size: a rigid shift of the nuclei shifts every electron rigidly
    shift = np.array([0.3, -0.2, 0.1])
    rigid = aiqmc_b200.correlated_samples(atoms, atoms + shift, torch.tensor(pos)).cpu().numpy()
    np.testing.assert_allclose(rigid.reshape(B, n, 3), pos.reshape(B, n, 3) + shift, rtol=1e-12, atol=1e-12)
    assert aiqmc_b200.correlated_samples(atoms, new_atoms, torch.zeros((0, 3 * n))).shape == (0, 3 * n)
