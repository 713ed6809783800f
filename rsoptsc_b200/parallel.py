"""Multi-GPU host plumbing: one process per GPU, no torch.

`init_comm` hands the NCCL unique id of rank 0 to every rank and creates the library's
communicator (rsoptsc_comm_init).  From then on the C ABI is SPMD: every rank calls the same entry
point with the same inputs and one computeM problem is split across the GPUs inside the library
(column blocks of the ADMM state, collective SVT of the large connected blocks).

The id travels over a caller-supplied `broadcast(bytes_or_None) -> bytes` (bench.py passes a
torch.distributed one because torchrun already gave it a process group), or by default over a
plain TCP rendezvous on MASTER_ADDR : MASTER_PORT + 1 (standard library only).

Independent problems (the NMF rank sweep k = 5..30 of config 4, the PAM sweep) are dealt to the
ranks round-robin with `sweep`, which needs a second, replica-style use of the GPUs: call it
WITHOUT an active library communicator (every rank then works on its own units).
"""
import ctypes as C
import os
import pickle
import socket
import struct
import time

import numpy as np

from . import _lib


def world():
    """(rank, local_rank, world_size) from the launcher's environment (torchrun / mpirun style)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def partition(units, rank, world_size):
    """Round-robin share of `units` for `rank` (deterministic, covers every unit exactly once)."""
    units = list(units)
    return [u for i, u in enumerate(units) if i % world_size == rank]


# ---- a minimal TCP store: rank 0 serves, the others connect (standard library only) -------------
def _recv_exact(sock, n):
    buf = b""
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("rendezvous peer closed the connection")
        buf += chunk
    return buf


def _send_msg(sock, obj):
    payload = pickle.dumps(obj)
    sock.sendall(struct.pack("!Q", len(payload)) + payload)


def _recv_msg(sock):
    (n,) = struct.unpack("!Q", _recv_exact(sock, 8))
    return pickle.loads(_recv_exact(sock, n))


class TcpGroup:
    """all_gather / broadcast of small python objects between the ranks of one job."""

    def __init__(self, rank, world_size, addr=None, port=None, timeout=120.0):
        self.rank, self.world_size = int(rank), int(world_size)
        self.addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        self.port = int(port if port is not None else int(os.environ.get("MASTER_PORT", 29500)) + 1)
        self.peers = []
        self.sock = None
        if self.world_size == 1:
            return
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((self.addr, self.port))
            srv.listen(self.world_size)
            srv.settimeout(timeout)
            conns = {}
            while len(conns) < self.world_size - 1:
                c, _ = srv.accept()
                c.settimeout(timeout)
                conns[_recv_msg(c)] = c
            srv.close()
            self.peers = [conns[r] for r in range(1, self.world_size)]
        else:
            deadline = time.time() + timeout
            while True:
                try:
                    s = socket.create_connection((self.addr, self.port), timeout=5.0)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            s.settimeout(timeout)
            _send_msg(s, self.rank)
            self.sock = s

    def all_gather(self, obj):
        if self.world_size == 1:
            return [obj]
        if self.rank == 0:
            parts = [obj] + [_recv_msg(c) for c in self.peers]
            for c in self.peers:
                _send_msg(c, parts)
            return parts
        _send_msg(self.sock, obj)
        return _recv_msg(self.sock)

    def broadcast(self, obj):
        return self.all_gather(obj if self.rank == 0 else None)[0]

    def close(self):
        for c in self.peers:
            c.close()
        if self.sock is not None:
            self.sock.close()
        self.peers, self.sock = [], None


# ---- the library communicator ---------------------------------------------------------------------
def init_comm(rank=None, world_size=None, device=None, broadcast=None):
    """Bind this process to `device` (default LOCAL_RANK) and create the library's NCCL communicator.
    `broadcast(payload_or_None) -> payload` must return rank 0's payload on every rank."""
    r, lr, ws = world()
    rank = r if rank is None else int(rank)
    world_size = ws if world_size is None else int(world_size)
    lib = _lib.init(lr if device is None else device)
    if world_size == 1:
        _lib.check(lib.rsoptsc_comm_init(1, 0, None))
        return lib
    buf = (C.c_uint8 * 128)()
    group = None
    if broadcast is None:
        group = TcpGroup(rank, world_size)
        broadcast = group.broadcast
    payload = None
    if rank == 0:
        _lib.check(lib.rsoptsc_comm_unique_id(buf))
        payload = bytes(buf)
    payload = broadcast(payload)
    if group is not None:
        group.close()
    if len(payload) != 128:
        raise _lib.RsoptscError("init_comm: the broadcast did not deliver the 128-byte NCCL unique id")
    buf = (C.c_uint8 * 128).from_buffer_copy(payload)
    _lib.check(lib.rsoptsc_comm_init(world_size, rank, buf))
    return lib


def finalize_comm():
    _lib.check(_lib.load().rsoptsc_comm_finalize())


def comm_info():
    """dict(nranks, rank, p2p, collectives, bytes) of the active library communicator."""
    lib = _lib.load()
    nr, rk, p2p, ncoll, nbytes = C.c_int32(1), C.c_int32(0), C.c_int32(0), C.c_int64(0), C.c_double(0)
    _lib.check(lib.rsoptsc_comm_info(C.byref(nr), C.byref(rk), C.byref(p2p), C.byref(ncoll), C.byref(nbytes)))
    return dict(nranks=nr.value, rank=rk.value, p2p=bool(p2p.value), collectives=ncoll.value, bytes=nbytes.value)


def shard_owner(idx, nranks, split_min=2048):
    """Host-only view of the multi-GPU split (rsoptsc_host_shard_owner): per cell the owning rank, its
    connected component and whether that component is cut across all ranks."""
    lib = _lib.load()
    ix = np.ascontiguousarray(idx, dtype=np.int32)
    n, K = ix.shape
    owner = np.full(n, -1, dtype=np.int32)
    comp = np.zeros(n, dtype=np.int32)
    split = np.zeros(n, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(_lib.c_int32_p)  # noqa: E731
    _lib.check(lib.rsoptsc_host_shard_owner(ip(ix), n, K, int(nranks), int(split_min), ip(owner), ip(comp), ip(split)))
    return dict(owner=owner, component=comp, split=split.astype(bool))


# ---- independent problems dealt to the ranks (replica-style use of the GPUs) ----------------------
def sweep(units, fn, group=None):
    """Run fn(unit) for this rank's share of `units`, return {unit: result} for ALL units.
    `group` is anything with all_gather(obj) -> list (a TcpGroup; tests pass a gloo-backed shim)."""
    rank, _, ws = world()
    if group is not None:
        rank, ws = group.rank, group.world_size
    local = {u: fn(u) for u in partition(units, rank, ws)}
    if ws == 1:
        return local
    own = group is None
    if own:
        group = TcpGroup(rank, ws)
    out = {}
    for part in group.all_gather(local):
        out.update(part)
    if own:
        group.close()
    return out


def nmf_rank_sweep(W, ranks=range(5, 31), group=None):
    """Config 4: ClusterCells for every rank k, the 26 independent NMF problems dealt to the GPUs."""
    from . import api

    def one(k):
        r = api.nmf_lee(W, k)
        return dict(labels=r["labels"], iters=r["iters"])
    return sweep(list(ranks), one, group)
