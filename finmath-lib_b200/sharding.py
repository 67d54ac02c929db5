"""Path sharding across GPUs.

Paths are independent: rank g of G owns the contiguous path block [g*P/G, (g+1)*P/G) of ONE logical simulation; its
MT19937 sub-stream starts at word 2*T*F*path_offset (jump-ahead), so the union over ranks is bit-identical to the
single-stream reference.  No path data ever moves between GPUs; the only exchange is the tiny reduction behind
getAverage / getVariance / getMin / getMax and the regression's 27 moments: per rank a double-double pair per sum,
all-gathered over torch.distributed (NCCL on GPUs, gloo in the CPU tests) and merged in rank order on every rank, so
all ranks hold the same bits.
"""
import numpy as np


def _two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def dd_merge(hi1, lo1, hi2, lo2):
    s, e = _two_sum(hi1, hi2)
    e += lo1 + lo2
    return _two_sum(s, e)


class ShardContext:
    def __init__(self, rank=0, world=1, group=None, device=None):
        self.rank, self.world, self.group, self.device = int(rank), int(world), group, device
        self.collectives = 0
        self._buffers = {}

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
        import torch
        import torch.distributed as dist
        self.collectives += 1
        n = len(values)
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
        if self.world == 1:
            return hi, lo
        g = self._all_gather([hi, lo])
        h, l = float(g[0, 0]), float(g[0, 1])
        for r in range(1, self.world):
            h, l = dd_merge(h, l, float(g[r, 0]), float(g[r, 1]))
        return h, l

    def sum_dd_many(self, his, los):
        """Element-wise sum over ranks of arrays of double-double pairs (regression moments: one message)."""
        his, los = np.asarray(his, dtype=np.float64), np.asarray(los, dtype=np.float64)
        if self.world == 1:
            return his, los
        g = self._all_gather(list(his.ravel()) + list(los.ravel()))
        n = his.size
        H, L = g[0, :n].copy(), g[0, n:].copy()
        for r in range(1, self.world):
            for i in range(n):
                H[i], L[i] = dd_merge(H[i], L[i], g[r, i], g[r, n + i])
        return H.reshape(his.shape), L.reshape(los.shape)

    def min(self, v):
        if self.world == 1:
            return v
        g = self._all_gather([v])[:, 0]
        return float(np.nan) if np.isnan(g).any() else float(g.min())

    def max(self, v):
        if self.world == 1:
            return v
        g = self._all_gather([v])[:, 0]
        return float(np.nan) if np.isnan(g).any() else float(g.max())

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
    return ShardContext(rank, world, None, device)
