"""world_size-2 gloo test of the data-parallel plumbing (the N>1 path of SURVEY 8(e)) on CPU."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from merlot_b200.train import DataParallel
    dp = DataParallel("gloo")
    g = torch.full((10,), float(rank + 1))
    dp.all_reduce_grads(g)
    x = torch.arange(6, dtype=torch.float32).reshape(3, 2) + 100 * rank
    try:
        allx = dp.all_gather_rows(x)
    except Exception:  # older gloo builds lack all_gather_into_tensor
        lst = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(lst, x)
        allx = torch.cat(lst, 0)
    # gradient of the gather: every rank contributes d_all; rank r receives sum over ranks of rows [r*n, (r+1)*n)
    d_all = torch.ones(world * 3, 2) * (rank + 1)
    try:
        mine = dp.reduce_scatter_rows(d_all)
    except Exception:
        dist.all_reduce(d_all)
        mine = d_all[rank * 3:(rank + 1) * 3]
    labels = torch.arange(3) + rank * 3  # model/modeling.py:519
    q.put((rank, g.tolist(), allx.tolist(), mine.tolist(), labels.tolist()))
    dist.destroy_process_group()


def test_gloo_world2_collectives():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    for rank, g, allx, mine, labels in res:
        assert g == [3.0] * 10  # sum over replicas; the mean's 1/world is folded into AdamW's grad_scale
        assert allx == [[0.0, 1.0], [2.0, 3.0], [4.0, 5.0], [100.0, 101.0], [102.0, 103.0], [104.0, 105.0]]
        assert mine == [[3.0, 3.0]] * 3
        assert labels == [rank * 3, rank * 3 + 1, rank * 3 + 2]
