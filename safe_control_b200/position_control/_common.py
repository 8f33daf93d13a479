"""Shared host logic of the single-agent shims."""
import numpy as np

from .._abi import STATUS_STR
from ..batched import HostContext

_ctx = {}


def host_ctx(device=0):
    """One staging context per device, created on first use (needs a CUDA device: no CPU path)."""
    if device not in _ctx:
        _ctx[device] = HostContext(device)
    return _ctx[device]


def obs_rows(obs, num_obs):
    """The reference accepts None, one obstacle as (n,) / (n,1), or rows (k, n) with n in {3, 5, 7}
    (tracking.py:277-291 pads to 7).  -> OBS [1, num_obs, 7] f64, nobs [1] i32 (-1 for None)."""
    OBS = np.zeros((1, max(num_obs, 1), 7))
    if obs is None:
        return OBS, np.array([-1], np.int32)
    a = np.asarray(obs, dtype=np.float64)
    if a.ndim == 3:                                # a list of (n, 1) columns (what Unicycle2D.agent_barrier expects)
        a = a.reshape(a.shape[0], -1)
    if a.ndim == 1 or (a.ndim == 2 and a.shape[1] == 1):
        a = a.reshape(1, -1)
    if a.shape[1] > 7:
        a = a[:, :7]
    k = min(a.shape[0], num_obs)
    OBS[0, :k, : a.shape[1]] = a[:k]
    return OBS, np.array([k], np.int32)


def status_string(code):
    return STATUS_STR.get(int(code), "solver_error")


class PinnedIO:
    """Page-locked (hence device-mapped) host buffers of ONE controller object, allocated at its first solve: with them
    the host-pointer C calls take the zero-copy path (the kernel reads X / u_ref / obstacles and writes u / status over
    PCIe directly, csrc/scb_api.cu) instead of seven staged cudaMemcpyAsync -- what the reference-side binding of
    INTEGRATION.md gets for free when it keeps its arrays in such buffers."""

    def __init__(self, **shapes):
        import torch
        self.buf = {}
        pin = torch.cuda.is_available()         # (buffers only: without a device the solve call itself raises)
        for name, (shape, dtype) in shapes.items():
            t = torch.empty(shape, dtype=dtype)
            self.buf[name] = (t.pin_memory() if pin else t).numpy()
        self._keep = None

    def __getitem__(self, k):
        return self.buf[k]

    def put(self, k, value):
        b = self.buf[k]
        b[...] = value
        return b
