"""Path sharding across GPUs.

Paths are independent: rank g of G owns the contiguous path block [g*P/G, (g+1)*P/G) of ONE logical simulation; its
MT19937 sub-stream starts at word 2*T*F*path_offset (jump-ahead), so the union over ranks is bit-identical to the
single-stream reference.  No path data ever moves between GPUs; the only exchange is the tiny reduction behind
getAverage / getVariance / getMin / getMax and the regression's 27 moments: per rank a double-double pair per sum,
all-gathered and merged in rank order on every rank, so all ranks hold the same bits.

Default on GPUs (FMB_TINY_COLLECTIVES=peer): the exchange happens INSIDE the native library, on its compute stream — every rank stores
its partials straight into the gather buffers of all ranks over NVLink (CUDA IPC mappings of one small buffer per rank) from one small
kernel, the partials are merged in rank order on the device, so a regression needs no host round trip and a getAverage is one
synchronisation.  =nccl does the same with an NCCL all-gather per exchange (the fallback when the devices have no peer access, and
across nodes).  The host-side variants remain as measured alternatives and for the CPU tests: FMB_TINY_COLLECTIVES=shm (a shared-memory
mailbox between the processes of one x86 node) and =torch (torch.distributed all-gather: gloo on CPU, NCCL with host staging on GPUs).
"""
import atexit
import os
import time

import numpy as np


def _two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def dd_merge(hi1, lo1, hi2, lo2):
    s, e = _two_sum(hi1, hi2)
    e += lo1 + lo2
    return _two_sum(s, e)


class _Mailbox:
    """All-gather of a few doubles between the processes of one node through POSIX shared memory.  Layout: per rank one cache line
    with a sequence number, then two message buffers (alternating, so that a fast rank that is already in the next exchange never
    overwrites what a slow rank still reads: a rank can only be two exchanges ahead after everybody has published the one in
    between, i.e. has finished reading the previous one)."""
    MAXN = 128                                             # doubles per message (the 27 + 27 regression moments fit)

    def __init__(self, rank, world, name, create):
        from multiprocessing import shared_memory
        self.rank, self.world, self.seq = rank, world, 0
        nbytes = world * 64 + 2 * world * self.MAXN * 8
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=nbytes)
        if not create:
            # only the creating rank owns (and unlinks) the segment; keep this process's resource tracker out of it
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.seqs = np.ndarray((world, 8), dtype=np.int64, buffer=self.shm.buf, offset=0)
        self.data = np.ndarray((2, world, self.MAXN), dtype=np.float64, buffer=self.shm.buf, offset=world * 64)
        if create:
            self.seqs[:] = 0
        self.owner = create

    def all_gather(self, values, timeout=120.0):
        n = len(values)
        s = self.seq + 1
        b = s & 1
        self.data[b, self.rank, :n] = values
        # publish after the payload.  Plain stores and loads through numpy views: correct only where the hardware keeps stores (and
        # loads) in program order, i.e. x86-64 (TSO) - from_environment() enables the mailbox on such hosts only; everywhere else
        # (aarch64: Grace) the exchange goes through the native communicator or torch.distributed.
        self.seqs[self.rank, 0] = s
        deadline = None
        for r in range(self.world):
            spins = 0
            while self.seqs[r, 0] < s:
                spins += 1
                if spins > 2000:                           # a peer is far behind (e.g. still computing): stop burning the core
                    if deadline is None:
                        deadline = time.time() + timeout
                    elif time.time() > deadline:
                        raise RuntimeError("finmath_b200: rank %d did not join a reduction within %.0f s" % (r, timeout))
                    time.sleep(0.00002)
        out = self.data[b, :, :n].copy()
        self.seq = s
        return out

    def close(self):
        try:
            self.seqs = self.data = None
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:
            pass


class ShardContext:
    def __init__(self, rank=0, world=1, group=None, device=None):
        self.rank, self.world, self.group, self.device = int(rank), int(world), group, device
        self.collectives = 0
        self._buffers = {}
        self._mailbox = None
        self.native_comm = False                             # True: the native library exchanges the partials itself (fmb_comm_init)
        self.peer_exchange = False                           # True: through peer memory over NVLink instead of NCCL all-gathers

    def use_native_comm(self):
        """Give the native library its own NCCL communicator (rank 0 creates the id, torch.distributed carries it to the others)."""
        import ctypes as C
        import torch.distributed as dist
        from . import native as nv
        lib = nv.load()
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            nv.check(lib.fmb_comm_unique_id(buf, 128))
        box = [buf.raw if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=self.group)
        nv.check(lib.fmb_comm_init(box[0], 128, self.rank, self.world))
        self.native_comm = True
        self.peer_exchange = False

    def use_peer_exchange(self):
        """Map the gather buffers of all ranks into every process (CUDA IPC) so that the native exchanges go through peer memory instead of
        NCCL.  All ranks of one node, 2..8 of them; returns False (and changes nothing, on any rank) if some rank cannot map a peer."""
        import ctypes as C
        import socket
        import torch.distributed as dist
        from . import native as nv
        lib = nv.load()
        if not self.native_comm or not 2 <= self.world <= 8:
            return False
        buf = C.create_string_buffer(64)
        ok = lib.fmb_comm_peer_handle(buf, 64) == nv.FMB_OK
        mine = (socket.gethostname(), buf.raw if ok else None)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        if len(set(h for h, _ in everyone)) != 1 or any(b is None for _, b in everyone):
            return False
        # two phases, so that either every rank switches or none does: map (can fail), agree, then switch
        rc = lib.fmb_comm_peer_open(b"".join(b for _, b in everyone), 64 * self.world)
        oks = [None] * self.world
        dist.all_gather_object(oks, rc == nv.FMB_OK, group=self.group)
        if not all(oks):
            if rc == nv.FMB_OK:
                nv.check(lib.fmb_comm_peer_open(None, 0))    # some other rank could not map its peers: everybody stays on NCCL
            return False
        self.peer_exchange = True
        return True

    def use_mailbox(self, name, create):
        """Switch the tiny all-gathers to the shared-memory mailbox (all ranks on one node)."""
        self._mailbox = _Mailbox(self.rank, self.world, name, create)
        atexit.register(self._mailbox.close)

    # ---- partition -----------------------------------------------------------------------------------------------
    def local_range(self, n_global):
        """[lo, hi) of this rank's contiguous path block."""
        if self.world == 1:
            return 0, n_global
        return (n_global * self.rank) // self.world, (n_global * (self.rank + 1)) // self.world

    def local_count(self, n_global):
        lo, hi = self.local_range(n_global)
        return hi - lo

    def global_count(self, n_local):
        if self.world == 1:
            return n_local
        return int(sum(self._all_gather([float(n_local)])[:, 0]))

    # ---- collectives ---------------------------------------------------------------------------------------------
    def _all_gather(self, values):
        """values: list of floats -> array [world][len(values)], identical on every rank.  One small NCCL all-gather over
        NVLink (gloo on CPU); buffers are cached per message length, the result comes back in a single device-to-host copy."""
        self.collectives += 1
        n = len(values)
        if self._mailbox is not None and n <= _Mailbox.MAXN:
            return self._mailbox.all_gather(values)
        import torch
        import torch.distributed as dist
        bufs = self._buffers.get(n)
        if bufs is None:
            on_gpu = self.device is not None
            dev = self.device if on_gpu else "cpu"
            src_host = torch.empty(n, dtype=torch.float64, pin_memory=on_gpu)
            out_host = torch.empty(self.world * n, dtype=torch.float64, pin_memory=on_gpu)
            bufs = (src_host, src_host.numpy(), torch.empty(n, dtype=torch.float64, device=dev),
                    torch.empty(self.world * n, dtype=torch.float64, device=dev), out_host, out_host.numpy())
            self._buffers[n] = bufs
        src_host, src_np, src, out, out_host, out_np = bufs
        src_np[:] = values
        src.copy_(src_host, non_blocking=True)
        dist.all_gather_into_tensor(out, src, group=self.group)
        out_host.copy_(out)                                  # synchronising device-to-host copy into the cached pinned buffer
        return out_np.reshape(self.world, n).copy()

    def sum_dd(self, hi, lo):
        if self.world == 1 or self.native_comm:              # (native communicator: fmb_rv_reduce already returned the sum over all shards)
            return hi, lo
        g = self._all_gather([hi, lo])
        h, l = float(g[0, 0]), float(g[0, 1])
        for r in range(1, self.world):
            h, l = dd_merge(h, l, float(g[r, 0]), float(g[r, 1]))
        return h, l

    def sum_dd_many(self, his, los, force=False):
        """Element-wise sum over ranks of arrays of double-double pairs (regression moments: one message).  force: the values are
        LOCAL partials even though a native communicator exists (fmb_regression_moments always returns local sums)."""
        his, los = np.asarray(his, dtype=np.float64), np.asarray(los, dtype=np.float64)
        if self.world == 1 or (self.native_comm and not force):
            return his, los
        g = self._all_gather(list(his.ravel()) + list(los.ravel()))
        n = his.size
        H, L = g[0, :n].copy(), g[0, n:].copy()
        for r in range(1, self.world):
            for i in range(n):
                H[i], L[i] = dd_merge(H[i], L[i], g[r, i], g[r, n + i])
        return H.reshape(his.shape), L.reshape(los.shape)

    def _extreme(self, v, has_data, pick):
        if self.world == 1 or self.native_comm:
            return v
        g = self._all_gather([v if has_data else 0.0, 1.0 if has_data else 0.0])
        vals = g[g[:, 1] != 0.0, 0]                          # shards that own no element do not take part (their NaN is "no data")
        if vals.size == 0:
            return float(np.nan)
        return float(np.nan) if np.isnan(vals).any() else float(pick(vals))

    def min(self, v, has_data=True):
        return self._extreme(v, has_data, np.min)

    def max(self, v, has_data=True):
        return self._extreme(v, has_data, np.max)

    def gather(self, local_array):
        """Concatenate the shards in rank order (getRealizations of the logical vector)."""
        if self.world == 1:
            return local_array
        import torch
        import torch.distributed as dist
        self.collectives += 1
        objs = [None] * self.world
        dist.all_gather_object(objs, np.ascontiguousarray(local_array), group=self.group)
        return np.concatenate(objs)

    def get_element(self, dv, i):
        if self.world == 1:
            return dv.get(i)
        # the owner rank reads it, everybody gets it
        counts = self._all_gather([float(dv.n)])[:, 0].astype(np.int64)
        start = int(counts[:self.rank].sum())
        v = dv.get(i - start) if start <= i < start + dv.n else 0.0
        return float(self._all_gather([v])[:, 0].sum())


LOCAL = ShardContext()


def from_environment(backend=None):
    """ShardContext for a torchrun launch (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); single process otherwise."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return LOCAL
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    device = None
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    shard = ShardContext(rank, world, None, device)
    from . import native as nv
    mode = os.environ.get("FMB_TINY_COLLECTIVES", "peer" if backend == "nccl" else "shm")
    if backend == "nccl":
        nv.init(local_rank)                                  # one process per GPU: the native library on this rank's device
    if mode in ("peer", "nccl") and backend == "nccl":
        shard.use_native_comm()
        if mode == "peer":
            shard.use_peer_exchange()                        # (stays on NCCL if the devices cannot map each other)
        return shard
    import platform
    if mode != "torch" and platform.machine() in ("x86_64", "AMD64"):
        # all ranks on one node (the launch contract of bench.py): host-resident partials go through shared memory
        import socket
        hosts = [None] * world
        dist.all_gather_object(hosts, socket.gethostname())
        if len(set(hosts)) == 1:
            name = "fmb_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid() if "TORCHELASTIC_RUN_ID" in os.environ else 0)
            names = [None] * world
            dist.all_gather_object(names, name + "_%d" % os.getpid() if rank == 0 else None)
            name = names[0]
            ok = True
            try:
                if rank == 0:
                    shard.use_mailbox(name, True)
            except Exception:
                ok = False
            dist.barrier()
            try:
                if rank != 0:
                    shard.use_mailbox(name, False)
            except Exception:
                ok = False
            oks = [None] * world
            dist.all_gather_object(oks, ok)
            if not all(oks):                                   # somebody could not map the segment: everybody uses torch.distributed
                if shard._mailbox is not None:
                    shard._mailbox.close()
                shard._mailbox = None
    return shard
