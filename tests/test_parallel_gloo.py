"""CPU (-m "not gpu"): the N>1 sharding plumbing (panfusion_b200/parallel.py) under gloo, world_size 2 — both
layouts a 2-rank job can take (CFG split and view split): slices, K|V all-gather order, output re-assembly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, layout, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from panfusion_b200.parallel import ViewParallel
        b, m, P, C = 2, 4, 3, 2
        par = ViewParallel(None, *layout)
        par.configure(b, m)
        bsl, vsl = par.slices(b, m)
        # global K|V tensor: value encodes (batch, view, pixel)
        full = torch.arange(b * m * P * C, dtype=torch.float32).reshape(b, m, P, C)
        loc = full[bsl, vsl].reshape(bsl.stop - bsl.start, -1, C).contiguous()
        got = par.gather_views(loc)
        want = full[bsl].reshape(bsl.stop - bsl.start, m * P, C)
        ok1 = torch.equal(got, want)
        sample = torch.arange(b * m * 5, dtype=torch.float32).reshape(b, m, 5)
        pano = torch.arange(b * 7, dtype=torch.float32).reshape(b, 1, 7)
        s, p = par.gather_outputs(sample[bsl, vsl].contiguous(), pano[bsl].contiguous(), b, m)
        ret[rank] = bool(ok1 and torch.equal(s, sample) and torch.equal(p, pano))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("layout", [(2, 1), (1, 2)])
def test_view_parallel_world2(layout):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, layout, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_pick_layout():
    from panfusion_b200.parallel import pick_layout
    assert pick_layout(1, 2, 8) == (1, 1)
    assert pick_layout(2, 2, 8) == (2, 1)   # pure CFG split: no collective inside EPPA
    assert pick_layout(4, 2, 8) == (2, 2)
    assert pick_layout(8, 2, 8) == (2, 4)
    assert pick_layout(8, 2, 20) == (2, 4)  # 20 icosahedron views, 5 per rank
    with pytest.raises(ValueError):
        pick_layout(8, 2, 6)
