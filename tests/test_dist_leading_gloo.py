"""N>1 host logic of the range-partitioned element-wise / reduction paths (sparse_b200/_dist.py: leading_splits,
leading_block, elemwise_leading, reduce_leading) on CPU: world_size-2 `gloo`, NumPy mock of the kernel layer; the
blocks must tile the single-process result exactly."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _mock_kernels

        _mock_kernels.install()
        import sparse_b200 as sp
        from sparse_b200 import _dist as DD

        rng = np.random.default_rng(77)  # same inputs on every rank
        a = sp.random((12, 5, 6), density=0.3, random_state=rng)
        b = sp.random((12, 5, 1), density=0.4, random_state=rng)   # broadcast on the trailing axis
        c = sp.random((1, 5, 6), density=0.5, random_state=rng)    # broadcast on the LEADING axis: replicated
        bounds = DD.leading_splits([a, b, c], world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        al, bl, cl = (DD.leading_block(x, r0, r1) for x in (a, b, c))
        res = {}
        full = (a + b) * c
        loc = DD.elemwise_leading(np.multiply, DD.elemwise_leading(np.add, al, bl), cl)
        res["elemwise"] = bool(np.array_equal(loc.todense(), full.todense()[r0:r1]) and loc.shape[0] == r1 - r0)
        # axis 0 kept: purely local
        s12 = DD.reduce_leading(al, np.add, axis=(1, 2))
        res["reduce_kept"] = bool(np.allclose(s12.todense(), a.todense().sum(axis=(1, 2))[r0:r1], rtol=1e-12))
        # axis 0 reduced: partials gathered and combined; identical on every rank
        for name, method, npf in (("sum0", np.add, np.sum), ("max0", np.maximum, np.max)):
            got = DD.reduce_leading(al, method, axis=0)
            res[name] = bool(np.allclose(got.todense(), npf(a.todense(), axis=0), rtol=1e-12) and got.shape == (5, 6))
        got = DD.reduce_leading(al, np.add, axis=(0, 2), keepdims=True)
        res["sum02_keepdims"] = bool(np.allclose(got.todense(), a.todense().sum(axis=(0, 2), keepdims=True), rtol=1e-12))
        got = DD.reduce_leading(al, np.add, axis=None)
        res["sum_all"] = bool(np.allclose(got.todense(), a.todense().sum(), rtol=1e-12) and got.shape == ())
        f = sp.COO(a.coords, a.data, shape=a.shape, fill_value=2.0, has_duplicates=False, sorted=True)
        fl = DD.leading_block(f, r0, r1)
        got = DD.reduce_leading(fl, np.maximum, axis=0)
        res["max0_fill"] = bool(np.array_equal(got.todense(), f.todense().max(axis=0)))
        try:
            DD.reduce_leading(fl, np.add, axis=0)
            res["sum0_fill_raises"] = False
        except ValueError:
            res["sum0_fill_raises"] = True
        q.put((rank, res, bounds))
    finally:
        dist.destroy_process_group()


def test_leading_partition_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, r, bounds in res:
        assert bounds[0] == 0 and bounds[-1] == 12 and 0 < bounds[1] < 12
        for k, ok in r.items():
            assert ok, f"rank {rank}: {k}"
