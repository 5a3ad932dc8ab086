"""Walker / parameter checkpoint wire format and the CSV log writer: the on-disk formats either side
of the walker hot path, so that a run can hand its wavefunction to the reference and back (DMC restarts from a VMC
checkpoint at DMC/main_dmc.py:65-72).

Mirrors AIQMCrelease3/checkpoint.py:13-70 -- `qmcjax_ckpt_{t:06d}.npz` written by np.savez with four members: `t`
(int), `data` (dataclasses.asdict(AINetData): positions / spins / atoms / charges), `params` and `opt_state`
(pytrees, pickled as 0-d object arrays) -- and AIQMCrelease3/utils/writers.py:7-46 (CSV with an iteration column).

Files written here hold plain numpy arrays, so the reference's `restore` (np.load(allow_pickle=True)) reads them
unchanged.  Files written by the reference hold pickled jax Arrays, which np.load can only rebuild with jax
importable; `restore` therefore walks the .npz itself and unpickles with a shim that turns jax's
`_reconstruct_array(fun, args, arr_state, aval_state)` records back into numpy arrays (jax is absent from this image:
the shim is exercised on emulated records, stated in the test)."""
from __future__ import annotations

import dataclasses
import datetime
import io
import os
import pickle
import zipfile
from typing import Any, Optional, Sequence

import numpy as np

from .api import AINetData


def _to_numpy_tree(t):
    if dataclasses.is_dataclass(t) and not isinstance(t, type):
        return _to_numpy_tree(dataclasses.asdict(t))
    if isinstance(t, dict):
        return {k: _to_numpy_tree(v) for k, v in t.items()}
    if isinstance(t, tuple) and hasattr(t, "_fields"):          # NamedTuple node (optax / kfac optimiser states)
        return type(t)(*(_to_numpy_tree(v) for v in t))
    if isinstance(t, (list, tuple)):
        return type(t)(_to_numpy_tree(v) for v in t)
    if hasattr(t, "detach"):                      # torch tensor (device or host)
        return t.detach().cpu().numpy()
    if t is None or isinstance(t, (int, float, complex, str, bool, np.generic)):
        return t
    return np.asarray(t)


def _readable(path: str) -> bool:
    """True if `path` is an intact .npz holding at least the step counter."""
    try:
        with zipfile.ZipFile(path) as z:
            return z.testzip() is None and 't.npy' in z.namelist()
    except (OSError, EOFError, zipfile.BadZipFile):
        return False


def find_last_checkpoint(ckpt_path: Optional[str] = None) -> Optional[str]:
    """Behaviour of checkpoint.py:13-25: the newest `qmcjax_ckpt*` file of the directory that can be opened; damaged
    files (an interrupted write) are skipped in favour of the next older one; None if there is nothing usable."""
    if not ckpt_path or not os.path.isdir(ckpt_path):
        return None
    names = sorted((n for n in os.listdir(ckpt_path) if 'qmcjax_ckpt' in n), reverse=True)
    return next((os.path.join(ckpt_path, n) for n in names if _readable(os.path.join(ckpt_path, n))), None)


def create_save_path(save_path: Optional[str]) -> str:
    """Behaviour of checkpoint.py:28-34: the given directory, or `./AInet_<timestamp>`, created on demand."""
    if not save_path:
        save_path = os.path.join(os.getcwd(), datetime.datetime.now().strftime('AInet_%Y_%m_%d_%H:%M:%S'))
    os.makedirs(save_path, exist_ok=True)
    return save_path


def get_restore_path(restore_path: Optional[str] = None) -> Optional[str]:
    return restore_path or None


def save(save_path: str, t: int, data: AINetData, params, opt_state=None) -> str:
    """checkpoint.py:46-61: same file name, member names and nesting; leaves become numpy arrays."""
    ckpt_filename = os.path.join(save_path, f'qmcjax_ckpt_{t:06d}.npz')
    def boxed(tree):
        # a pytree member is ONE pickled object (what np.savez makes of a dict); sequences (optax / kfac states are
        # tuples of NamedTuples) must be boxed explicitly or numpy tries to build a ragged array out of them
        cell = np.empty((), dtype=object)
        cell[()] = _to_numpy_tree(tree)
        return cell
    with open(ckpt_filename, 'wb') as f:
        np.savez(f, t=t, data=boxed(data), params=boxed(params), opt_state=boxed(opt_state))
    return ckpt_filename


def _reconstruct_jax_array(fun, args, arr_state, aval_state=None):
    """numpy side of jax._src.array._reconstruct_array: rebuild the host copy, drop the device placement."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


class _Unpickler(pickle.Unpickler):
    """Plain pickle, except that jax Array records come back as numpy arrays and nothing else from jax is needed."""

    def find_class(self, module, name):
        if module.startswith("jax") or module.startswith("jaxlib"):
            if name == "_reconstruct_array":
                return _reconstruct_jax_array
            raise pickle.UnpicklingError(f"checkpoint references {module}.{name}: only jax Arrays are supported")
        return super().find_class(module, name)


def _load_member(z: zipfile.ZipFile, name: str):
    with z.open(name) as fh:
        buf = io.BytesIO(fh.read())
    version = np.lib.format.read_magic(buf)
    shape, _, dtype = (np.lib.format.read_array_header_1_0(buf) if version == (1, 0)
                       else np.lib.format.read_array_header_2_0(buf))
    if dtype.hasobject:
        obj = _Unpickler(buf).load()
        return obj.item() if isinstance(obj, np.ndarray) and obj.shape == () else obj
    buf.seek(0)
    arr = np.load(buf, allow_pickle=False)
    return arr.item() if arr.shape == () else arr


def restore(restore_filename: str, batch_size: Optional[int] = None):
    """checkpoint.py:64-70 -> (t + 1, AINetData, params, opt_state); `batch_size`, if given, must match the walkers
    stored (the reference ignores the argument)."""
    with zipfile.ZipFile(restore_filename) as z:
        t = int(_load_member(z, 't.npy')) + 1
        data = AINetData(**_load_member(z, 'data.npy'))
        params = _load_member(z, 'params.npy')
        opt_state = _load_member(z, 'opt_state.npy') if 'opt_state.npy' in z.namelist() else None
    if batch_size is not None and np.asarray(data.positions).reshape(-1, np.asarray(data.positions).shape[-1]).shape[0] != batch_size:
        raise ValueError(f"checkpoint holds {np.asarray(data.positions).shape} positions, batch_size {batch_size} requested")
    return t, data, params, opt_state


class Writer:
    """CSV training log with the on-disk layout of utils/writers.py:7-46: a header row `<iteration_key>,<schema...>`,
    one row per write(t, **values), empty cells for values not given, ValueError for a key outside the schema.
    Usable as a context manager."""

    def __init__(self, name: str, schema: Sequence[str], directory: str = 'logs/', iteration_key: Optional[str] = 't',
                 log: bool = False):
        self.columns = tuple(schema)
        self.iteration_key = iteration_key
        self.echo = log
        os.makedirs(directory, exist_ok=True)
        self.path = os.path.join(directory, f'{name}.csv')
        self._fh = None

    def _emit(self, cells):
        self._fh.write(','.join(cells) + '\n')

    def __enter__(self):
        self._fh = open(self.path, 'w', encoding='UTF-8')
        self._emit(([self.iteration_key] if self.iteration_key else []) + list(self.columns))
        return self

    def write(self, t: int, **values: Any):
        unknown = [k for k in values if k not in self.columns]
        if unknown:
            raise ValueError(f'Not a recognized key for writer: {unknown[0]}')
        cells = [str(values[c]) if c in values else '' for c in self.columns]
        self._emit(([str(t)] if self.iteration_key else []) + cells)
        if self.echo:
            print(f'Iteration {t}: {values}')

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self._fh is not None:
            self._fh.close()
            self._fh = None
        return False
