"""Import alias: the product package lives in the directory the task names
(`ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200/`), which is not a
valid Python identifier; this shim loads it under the name `aiqmc_b200`."""
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_long = os.path.join(_root, "ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_long, "__init__.py"),
                                               submodule_search_locations=[_long])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
