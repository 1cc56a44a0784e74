"""Symmetric (peer-mapped) device buffers for in-kernel NVLink communication.

Bootstrap only — outside every timed region: buffers are allocated with
``torch.distributed._symmetric_memory`` (CUDA VMM + fabric/fd handle exchange; yields peer
pointers and, where the NVSwitch supports it, a multicast pointer). If that backend is
unavailable in the container, a CUDA-IPC fallback (``cudaIpcGetMemHandle`` exchanged through
the process group) provides the peer pointers without multicast. With one rank the buffer is
an ordinary allocation.

The reference's "transport" is a pickle file per client on local disk
(FLPyfhelin.py:225-237, :308-309, :374).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .. import _ext


class SymmetricBuffer:
    """``numel`` int64 words (``dtype`` selectable) visible to every rank of ``group``."""

    def __init__(self, numel: int, dtype: torch.dtype = torch.int64, device: Optional[torch.device] = None,
                 group: Optional[dist.ProcessGroup] = None, backend: str = "auto"):
        self.ops = _ext.ops()
        self.numel = int(numel)
        self.dtype = dtype
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.mc_ptr = 0
        self.backend = "local"
        self._keep = []
        if self.world == 1:
            self.tensor = torch.zeros(self.numel, dtype=dtype, device=self.device)
            self.ptrs = [self.tensor.data_ptr()]
            return
        errors = []
        if backend in ("auto", "symm_mem"):
            try:
                self._init_symm_mem()
                return
            except Exception as e:  # noqa: BLE001 - fall back to IPC, report both on failure
                errors.append(f"symm_mem: {type(e).__name__}: {e}")
                if backend == "symm_mem":
                    raise
        try:
            self._init_ipc()
        except Exception as e:  # noqa: BLE001
            errors.append(f"cuda_ipc: {type(e).__name__}: {e}")
            raise RuntimeError("no peer-memory backend available: " + " | ".join(errors)) from e
        if errors and self.rank == 0 and os.environ.get("HEFL_VERBOSE"):
            print("[hefl] symmetric memory fell back to CUDA IPC:", errors[0])

    # -- torch symmetric memory ---------------------------------------------------------
    def _init_symm_mem(self) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        grp = self.group if self.group is not None else dist.group.WORLD
        t = symm_mem.empty(self.numel, dtype=self.dtype, device=self.device)
        hdl = symm_mem.rendezvous(t, grp)
        t.zero_()
        self.tensor = t
        self.ptrs = [int(p) for p in hdl.buffer_ptrs]
        try:
            self.mc_ptr = int(hdl.multicast_ptr) if hdl.has_multicast_support(
                self.device.type, self.device.index or 0) else 0
        except Exception:  # noqa: BLE001
            try:
                self.mc_ptr = int(hdl.multicast_ptr)
            except Exception:  # noqa: BLE001
                self.mc_ptr = 0
        self._keep.append(hdl)
        self.backend = "symm_mem"
        torch.cuda.synchronize(self.device)
        dist.barrier(grp)

    # -- CUDA IPC fallback ------------------------------------------------------------------
    def _init_ipc(self) -> None:
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        nbytes = self.numel * torch.empty((), dtype=self.dtype).element_size()
        raw = self.ops.ipc_alloc(nbytes, dev)
        self._keep.append(raw)
        self.tensor = raw.view(self.dtype)
        handle = self.ops.ipc_get_handle(raw)
        gathered: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(gathered, bytes(handle.numpy().tobytes()), group=self.group)
        self.ptrs = []
        for r in range(self.world):
            if r == self.rank:
                self.ptrs.append(raw.data_ptr())
            else:
                h = torch.frombuffer(bytearray(gathered[r]), dtype=torch.uint8)
                self.ptrs.append(int(self.ops.ipc_open_handle(h, dev)))
        self.backend = "cuda_ipc"
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    def peer_view(self, rank: int) -> torch.Tensor:
        """Debug/test helper: peer ``rank``'s buffer as an int64 tensor on this device."""
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        nwords = self.numel * torch.empty((), dtype=self.dtype).element_size() // 8
        return self.ops.tensor_from_ptr(self.ptrs[rank], nwords, dev)


def probe() -> dict:
    """Report the peer-memory capabilities of this box (first thing to run on a new machine)."""
    info = {"cuda": torch.cuda.is_available()}
    if not torch.cuda.is_available():
        return info
    n = torch.cuda.device_count()
    info["devices"] = n
    info["name"] = torch.cuda.get_device_name(0)
    info["p2p"] = [[bool(torch.cuda.can_device_access_peer(i, j)) if i != j else True
                    for j in range(n)] for i in range(n)]
    try:
        import torch.distributed._symmetric_memory as symm_mem

        info["symm_mem_backend"] = str(symm_mem.get_backend(torch.device("cuda", 0)))
    except Exception as e:  # noqa: BLE001
        info["symm_mem_backend"] = f"error: {e}"
    return info
