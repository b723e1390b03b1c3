"""Result arrays in host memory that every rank of ONE node can write: the gather of a sharded movie
(generate_hologram_sequence.sequence_holograms) without a hop through rank 0's device.

The reference writes one ``.npy`` per frame from a single process (generate_hologram_sequence.py:19-26); the sharded
driver returns the whole movie on rank 0.  With a device-to-device gather every frame crosses rank 0's PCIe link
(1024 float64 frames at the SLM shape: 6 GiB over one link); here rank 0 creates a POSIX shared-memory segment, every
rank maps it, page-locks the part it fills and reads its own frames back into it over its OWN link.  Segments are kept
per process and used again by later calls of the same size once the array handed out before is gone (page-locking
hundreds of megabytes costs as much as the movie itself).
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import weakref
from typing import List, Optional

import numpy as np

_POOL: List[dict] = []            # this process's segments: name, shm, nbytes, owner (creator), holder (weakref), locked


def same_node(dist) -> bool:
    """All ranks of the default group on this host?  (torchrun sets LOCAL_WORLD_SIZE; without it: compare host names.)"""
    world = dist.get_world_size()
    lws = os.environ.get("LOCAL_WORLD_SIZE")
    if lws is not None:
        return int(lws) == world
    import socket
    names = [None] * world
    dist.all_gather_object(names, socket.gethostname())
    return len(set(names)) == 1


def _free_entry(nbytes: int, taken=()) -> Optional[dict]:
    for e in _POOL:
        if e["owner"] and e["nbytes"] == nbytes and (e["holder"] is None or e["holder"]() is None) and not any(e is t for t in taken):
            return e
    return None


def _attach(name: str, nbytes: int) -> dict:
    from multiprocessing import resource_tracker, shared_memory
    for e in _POOL:
        if e["name"] == name:
            return e
    shm = shared_memory.SharedMemory(name=name)
    try:                                              # the creator unlinks it; this process only maps it
        resource_tracker.unregister(shm._name, "shared_memory")
    except Exception:
        pass
    e = {"name": name, "shm": shm, "nbytes": nbytes, "owner": False, "holder": None, "locked": set()}
    _POOL.append(e)
    return e


def shared_results(dist, specs, my_rows, page_lock=True):
    """Collective over the default group.  ``specs``: [(shape, dtype), ...].  Returns one array per spec backed by
    memory shared by all ranks (rank 0 created or re-used the segments; ONE broadcast carries their names), the rows
    ``my_rows = (lo, hi)`` of each page-locked in this process -- or None on every rank when rank 0 could not create
    them (no /dev/shm, not enough room): the caller then gathers through the devices."""
    from multiprocessing import shared_memory
    sizes = [int(np.prod(shape)) * np.dtype(dtype).itemsize for shape, dtype in specs]
    rank = dist.get_rank()
    entries, msg = [], [None]
    if rank == 0 and all(n > 0 for n in sizes):
        try:
            for n in sizes:
                e = _free_entry(n, entries)
                if e is None:
                    # A tmpfs that is too small must fail HERE, not with SIGBUS in a copy.  (Not posix_fallocate: the pages
                    # should be first touched -- page-locked -- by the rank that fills them, on ITS memory node.)
                    st = os.statvfs("/dev/shm")
                    if n > 0.8 * st.f_bavail * st.f_frsize:
                        raise OSError("not enough room in /dev/shm")
                    shm = shared_memory.SharedMemory(create=True, size=n)
                    e = {"name": shm.name, "shm": shm, "nbytes": n, "owner": True, "holder": None, "locked": set()}
                    _POOL.append(e)
                entries.append(e)
            msg = [{"use": [e["name"] for e in entries], "drop": _surplus(entries)}]
        except Exception:
            entries, msg = [], [None]
    dist.broadcast_object_list(msg, src=0)
    if msg[0] is None:
        return None
    for name in msg[0]["drop"]:
        _drop(name)
    if rank != 0:
        entries = [_attach(name, n) for name, n in zip(msg[0]["use"], sizes)]   # (same node: what rank 0 created can be mapped)
    out = []
    lo, hi = my_rows
    for e, (shape, dtype) in zip(entries, specs):
        arr = np.ndarray(tuple(shape), dtype=dtype, buffer=e["shm"].buf)
        if rank == 0:
            e["holder"] = weakref.ref(arr)
        if page_lock and hi > lo:
            _page_lock(e, arr[lo:hi])
        out.append(arr)
    return out


def _surplus(in_use) -> List[str]:
    """rank 0: names of unused segments to give back once the pool holds more than SLM_SHARED_RESULT_BYTES (16 GiB)"""
    limit = int(os.environ.get("SLM_SHARED_RESULT_BYTES", 16 << 30))
    total = sum(e["nbytes"] for e in _POOL)
    names = []
    for e in _POOL:
        if total <= limit:
            break
        if e["owner"] and not any(e is u for u in in_use) and (e["holder"] is None or e["holder"]() is None):
            names.append(e["name"])
            total -= e["nbytes"]
    return names


def _drop(name: str) -> None:
    for i, e in enumerate(_POOL):
        if e["name"] == name:
            _release(e)
            del _POOL[i]
            return


def _release(e: dict) -> None:
    try:
        from . import _ffi
        lib = _ffi.load()
        for a, _ in e["locked"]:
            lib.slm_host_unregister(C.c_void_p(a))
    except Exception:
        pass
    try:
        e["shm"].close()
    except Exception:                                 # (an array handed out is still alive: the mapping stays until exit)
        pass
    if e["owner"]:
        try:
            e["shm"].unlink()
        except Exception:
            pass


def _page_lock(entry: dict, part: np.ndarray) -> None:
    """cudaHostRegister the pages under ``part`` (once per segment and range); failure leaves ordinary memory, which
    the read-back handles too (staged by the driver)."""
    page = 4096
    a = part.ctypes.data // page * page
    b = -(-(part.ctypes.data + part.nbytes) // page) * page
    if (a, b) in entry["locked"]:
        return
    try:
        from . import _ffi
        lib = _ffi.load()
        if lib.slm_host_register(C.c_void_p(a), b - a) == 0:
            entry["locked"].add((a, b))
    except Exception:
        pass


def _cleanup():
    for e in _POOL:
        _release(e)
    _POOL.clear()


atexit.register(_cleanup)
