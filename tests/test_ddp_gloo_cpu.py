"""N>1 path on CPU: two processes over gloo run the data-parallel step the way bench.py does on NCCL
(batch sharded by image, no forward communication, DDP gradient all-reduce, max-over-ranks timing) and
must reproduce the single-process gradients of the full batch.  The native ops are replaced by their
CPU stand-ins (tests/test_host_logic_cpu.py::install_cpu_ops); what is under test is the sharding /
collective harness, not the kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _loss(out):
    return out["pred_masks"].square().mean() + out["pred_logits"].square().mean()


def _build(setattr_fn=setattr):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here, os.path.join(here, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cases
    from oracle import torch_oracle as O
    from test_host_logic_cpu import build_decoder, install_cpu_ops
    from test_oracle_vs_golden import decoder_template
    install_cpu_ops(setattr_fn)
    dec = build_decoder()
    dec.load_state_dict(O.seeded_state_dict(decoder_template(), seed=51))
    x, mf = cases.decoder_inputs(B=4, seed=77)
    return dec, x, mf


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    dec, x, mf = _build()
    ddp = torch.nn.parallel.DistributedDataParallel(dec)
    per = x[0].shape[0] // world
    sl = slice(rank * per, (rank + 1) * per)
    out = ddp([t[sl] for t in x], mf[sl])
    _loss(out).backward()
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)           # the bench's max-over-ranks reduction
    grads = {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}
    # the flat gradient all-reduce bench.py uses after a CUDA-graph replay must equal DDP's bucketed one
    from mp_former_b200 import graphs
    for p in dec.parameters():
        p.grad = None
    _loss(dec([t_[sl] for t_ in x], mf[sl])).backward()
    graphs.allreduce_gradients(list(dec.parameters()), world)
    flat_ok = all((p.grad - grads[n]).abs().max().item() <= 1e-5 * max(1e-6, grads[n].abs().max().item())
                  for n, p in dec.named_parameters() if p.grad is not None)
    if rank == 0:
        ret["grads"] = grads
        ret["max"] = t.item()
        ret["flat_ok"] = flat_ok
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(monkeypatch):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["max"] == float(world)
    assert ret["flat_ok"]
    dec, x, mf = _build(monkeypatch.setattr)      # patches are undone after the test
    # DDP averages gradients over ranks; each rank's loss is the mean over its shard
    per = x[0].shape[0] // world
    total = 0
    for r in range(world):
        sl = slice(r * per, (r + 1) * per)
        total = total + _loss(dec([t[sl] for t in x], mf[sl])) / world
    total.backward()
    got = ret["grads"]
    n = 0
    for name, p in dec.named_parameters():
        if p.grad is None:
            continue
        scale = max(1e-6, p.grad.abs().max().item())
        assert (got[name] - p.grad).abs().max().item() / scale < 1e-4, name
        n += 1
    assert n > 40
