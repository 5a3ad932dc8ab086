"""bench.py host-side contract pieces that need no GPU: the bounded CPU sample of every workload and the shared config."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cpu_sample_is_bounded_for_every_workload():
    """The oracle's ccECP energy costs 50*N*A psi evaluations per walker: the CPU leg must shrink with the system
    (2,048 benzene walkers once ran a host out of memory).  Budget: <= ~2e6 single-electron-move evaluations of the
    network times its size per step."""
    b = _bench()
    from aiqmc_b200 import workloads as W
    for head in ("c_ecp", "c_ae", "n2", "c6h6", "dmc"):
        walkers, steps = b.cpu_sample(head, 0)
        sysd = W.SYSTEMS[head if head != "dmc" else "c_ecp"]
        n, a = len(sysd["spins"]), len(sysd["charges"])
        evals = walkers * (50 * n * a if head != "c_ae" else 6 * n) * n * n       # ~ evaluations x network size
        assert 1 <= walkers <= 2048 and 1 <= steps <= 5
        assert evals <= 2.0e8, (head, walkers, evals)
        assert b.cpu_sample(head, 7)[0] == 7                                       # explicit --cpu-walkers wins


def test_config_is_identical_for_both_arms():
    b = _bench()
    c1 = b.make_config("label", "c_ecp", 65536, 8)
    c2 = b.make_config("label", "c_ecp", 65536, 8)
    assert c1 == c2 and c1["global_walkers"] == 8 * 65536 and c1["workload_key"] == "c_ecp"
