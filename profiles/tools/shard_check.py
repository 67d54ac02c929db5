"""Multi-rank checks of the sharded reductions on real GPUs (torchrun, one process per GPU): every statistic of a sharded vector against numpy
on the whole vector — including vectors so short that some ranks own nothing — through whichever exchange FMB_TINY_COLLECTIVES selects
(peer memory by default, nccl, shm, torch).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 profiles/tools/shard_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
shard = pkg.from_environment()
RV = pkg.RandomVariableCuda
rng = np.random.default_rng(7)
failures = 0


def check(name, got, want, tol=0.0):
    global failures
    ok = (got == want) or (np.isnan(got) and np.isnan(want)) or abs(got - want) <= tol * max(1.0, abs(want))
    if not ok:
        failures += 1
        print("rank %d FAIL %s: got %r want %r" % (shard.rank, name, got, want), flush=True)


for n in (1, 2, 3, shard.world - 1, shard.world, shard.world + 1, 1000, 100_003):
    if n < 1:
        continue
    x = rng.standard_normal(n)
    w = rng.uniform(0.5, 1.5, n) / n
    v, ww = RV(0.0, x, shard), RV(0.0, w, shard)
    check("average n=%d" % n, v.getAverage(), float(np.sum(x.astype(np.longdouble)) / n), 1e-15)
    # (the reference divides the weighted sum by the number of paths as well: RandomVariableFromDoubleArray.java:296-318)
    check("weighted average n=%d" % n, v.getAverage(ww), float(np.sum((x * w).astype(np.longdouble)) / n), 1e-14)
    check("min n=%d" % n, v.getMin(), float(np.min(x)))
    check("max n=%d" % n, v.getMax(), float(np.max(x)))
    if n > 1:
        check("variance n=%d" % n, v.getVariance(), float(np.mean((x - np.mean(x)) ** 2)), 1e-13)
    check("median n=%d" % n, v.getQuantile(0.5), float(np.sort(x)[min(max(int(np.floor((n + 1) * 0.5 - 1 + 0.5)), 0), n - 1)]))
    check("sum of squares n=%d" % n, v.squared().add(1.0).getAverage(), float(np.mean(x * x + 1.0)), 1e-14)
# regression on sharded vectors (also with empty shards): coefficients of y = 2 + 3 b against the basis (1, b, b^2)
for n in (shard.world - 1 if shard.world > 1 else 1, 5, 50_001):
    n = max(n, 4)
    b = rng.standard_normal(n)
    y = 2.0 + 3.0 * b
    B, Y = RV(1.0, b, shard), RV(1.0, y, shard)
    est = pkg.MonteCarloConditionalExpectationRegression([B.mult(0.0).add(1.0), B, B.squared()])
    ce = Y.getConditionalExpectation(est)
    check("regression residual n=%d" % n, ce.sub(Y).squared().getAverage(), 0.0, 1e-18)
    coef = est.lastParameters
    check("regression intercept n=%d" % n, float(coef[0]), 2.0, 1e-9)
    check("regression slope n=%d" % n, float(coef[1]), 3.0, 1e-9)
import torch.distributed as dist  # noqa: E402
total = [None] * shard.world
dist.all_gather_object(total, failures)
if shard.rank == 0:
    print("shard_check: world %d, exchange %s, failures per rank %s" % (
        shard.world, "peer" if getattr(shard, "peer_exchange", False) else ("nccl" if shard.native_comm else "host"), total), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if sum(total) else 0)
