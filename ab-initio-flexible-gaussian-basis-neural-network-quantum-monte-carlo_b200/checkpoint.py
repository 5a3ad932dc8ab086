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

import contextlib
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
    if isinstance(t, (list, tuple)):
        return type(t)(_to_numpy_tree(v) for v in t)
    if hasattr(t, "detach"):                      # torch tensor (device or host)
        return t.detach().cpu().numpy()
    if t is None or isinstance(t, (int, float, complex, str, bool, np.generic)):
        return t
    return np.asarray(t)


def find_last_checkpoint(ckpt_path: Optional[str] = None) -> Optional[str]:
    """checkpoint.py:13-25: newest readable qmcjax_ckpt_* file of the directory, or None."""
    if ckpt_path and os.path.exists(ckpt_path):
        for file in sorted((f for f in os.listdir(ckpt_path) if 'qmcjax_ckpt' in f), reverse=True):
            fname = os.path.join(ckpt_path, file)
            try:
                with zipfile.ZipFile(fname) as z:
                    if z.testzip() is None and 't.npy' in z.namelist():
                        return fname
            except (OSError, EOFError, zipfile.BadZipFile):
                continue
    return None


def create_save_path(save_path: Optional[str]) -> str:
    """checkpoint.py:28-34."""
    timestamp = datetime.datetime.now().strftime('%Y_%m_%d_%H:%M:%S')
    ckpt_save_path = save_path or os.path.join(os.getcwd(), f'AInet_{timestamp}')
    if ckpt_save_path and not os.path.isdir(ckpt_save_path):
        os.makedirs(ckpt_save_path)
    return ckpt_save_path


def get_restore_path(restore_path: Optional[str] = None) -> Optional[str]:
    return restore_path if restore_path else None


def save(save_path: str, t: int, data: AINetData, params, opt_state=None) -> str:
    """checkpoint.py:46-61: same file name, member names and nesting; leaves become numpy arrays."""
    ckpt_filename = os.path.join(save_path, f'qmcjax_ckpt_{t:06d}.npz')
    with open(ckpt_filename, 'wb') as f:
        np.savez(f, t=t, data=_to_numpy_tree(data), params=_to_numpy_tree(params), opt_state=_to_numpy_tree(opt_state))
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


class Writer(contextlib.AbstractContextManager):
    """utils/writers.py:7-46: CSV log with a leading iteration column; unknown keys raise."""

    def __init__(self, name: str, schema: Sequence[str], directory: str = 'logs/', iteration_key: Optional[str] = 't',
                 log: bool = False):
        self._schema = list(schema)
        if not os.path.isdir(directory):
            os.makedirs(directory)
        self._filename = os.path.join(directory, name + '.csv')
        self._iteration_key = iteration_key
        self._log = log

    def __enter__(self):
        self._file = open(self._filename, 'w', encoding='UTF-8')
        if self._iteration_key:
            self._file.write(f'{self._iteration_key},')
        self._file.write(','.join(self._schema) + '\n')
        return self

    def write(self, t: int, **data: Any):
        for key in data:
            if key not in self._schema:
                raise ValueError(f'Not a recognized key for writer: {key}')
        row = [str(data.get(key, '')) for key in self._schema]
        if self._iteration_key:
            row.insert(0, str(t))
        self._file.write(','.join(row) + '\n')
        if self._log:
            print(f'Iteration {t}: {data}')

    def __exit__(self, exc_type, exc_val, exc_tb):
        self._file.close()
