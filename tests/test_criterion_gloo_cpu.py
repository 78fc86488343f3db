"""N>1 path of the criterion on CPU: two gloo ranks, one image each with different numbers of targets; the device
criterion's mask-count all-reduce (kept as a tensor, no .item()) must normalise both ranks' losses by the global
average (ref criterion.py:231-237), checked against the oracle told the global count.  Kernels emulated in host memory
as in tests/test_criterion_host_logic_cpu.py -- the collective / normalisation logic is under test, not the kernels."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import test_criterion_host_logic_cpu as T
    from make_golden_criterion import CFG, inputs
    from mp_former_b200 import _lib, native
    from oracle import criterion_oracle as CO
    _lib.require_cuda = lambda t, name: None                      # this process only
    native.point_sample_rows = T._emu_sample
    native.point_sample_rows_bwd = T._emu_sample_bwd
    native.topk_gather_rows = T._emu_topk_gather
    native.MaskLossRows = T._EmuMaskLossRows
    outputs, targets = inputs(with_dn=True)                       # 2 images: 3 and 5 targets

    def shard(x):
        if isinstance(x, torch.Tensor):
            return x[rank:rank + 1]
        if isinstance(x, list):
            return [shard(v) for v in x]
        if isinstance(x, dict):
            return {k: (v if k == "dn_args" else shard(v)) for k, v in x.items()}
        return x

    o, t = shard(outputs), targets[rank:rank + 1]
    crit = T._criterion().train(True)
    torch.manual_seed(40 + rank)
    got = crit(o, t)
    torch.manual_seed(40 + rank)
    ref = CO.set_criterion(o, t, losses=["labels", "masks"], training=True, world_size=world,
                           global_num_masks=sum(len(x["labels"]) for x in targets), **CFG)
    ok = sorted(got) == sorted(ref) and all(torch.allclose(got[k], ref[k], rtol=1e-5, atol=1e-6) for k in ref)
    # and it differs from the purely local normalisation (3 or 5 instead of 4)
    torch.manual_seed(40 + rank)
    local = CO.set_criterion(o, t, losses=["labels", "masks"], training=True, **CFG)
    differs = not torch.allclose(got["loss_mask"], local["loss_mask"], rtol=1e-3)
    ret[rank] = (bool(ok), bool(differs))
    dist.barrier()
    dist.destroy_process_group()


def test_criterion_two_rank_gloo_normalisation():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: (True, True), 1: (True, True)}
