"""FlatTrainer host logic on CPU: equivalence with the reference's step (zero_grad / backward / clip_grad_norm_ / Adam,
train.py:195-204) and, with world_size 2 over gloo, with a single-process step on the concatenated batch."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from domain_specific_image_compression_b200.trainer import FlatTrainer


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.b = torch.nn.Conv2d(8, 4, 3, padding=1)
        self.gamma = torch.nn.Parameter(torch.eye(4))      # dead parameter, like GDN's CxC gamma: never receives a gradient

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _loss(m, x):
    return (m(x) - 0.3).square().mean() * 50.0


def test_flat_trainer_equals_reference_step():
    torch.manual_seed(0)
    m = Toy()
    ref = copy.deepcopy(m)
    tr = FlatTrainer(m, lr=1e-2, grad_clip=1.0, fused=False, exclude=["gamma"])
    assert "gamma" not in tr.names and tr.flat.numel() == sum(p.numel() for n, p in ref.named_parameters() if n != "gamma")
    opt = torch.optim.Adam([p for n, p in ref.named_parameters() if n != "gamma"], lr=1e-2)
    x = torch.rand(4, 3, 8, 8)
    for _ in range(4):
        tr.step(lambda: _loss(m, x))
        opt.zero_grad(set_to_none=True)
        _loss(ref, x).backward()
        torch.nn.utils.clip_grad_norm_([p for n, p in ref.named_parameters() if n != "gamma"], 1.0)
        opt.step()
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p, q, atol=1e-7), n
    # parameters are views into the flat buffer: an optimizer step on the buffer is visible in the module
    assert m.a.weight.data_ptr() == tr.flat.data_ptr()


def test_buckets_partition_the_flat_buffer_and_fire_in_backward_order():
    """Reverse-order buckets: contiguous, disjoint, covering; with a small cap every layer is its own bucket and the hooks start
    the later layer's reduction first (its gradients exist first).  One big bucket gives the same update."""
    torch.manual_seed(1)
    m1, m2 = Toy(), Toy()
    m2.load_state_dict(m1.state_dict())
    small = FlatTrainer(m1, lr=1e-2, fused=False, exclude=["gamma"], bucket_bytes=8)
    big = FlatTrainer(m2, lr=1e-2, fused=False, exclude=["gamma"], bucket_bytes=1 << 30)
    assert len(big.buckets) == 1 and len(small.buckets) == 4
    spans = sorted((lo, hi) for lo, hi, _, _ in small.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == small.flat.numel() and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    x = torch.rand(2, 3, 8, 8)
    for _ in range(2):
        small.step(lambda: _loss(m1, x))
        big.step(lambda: _loss(m2, x))
    assert torch.equal(small.flat.detach(), big.flat.detach())
    first_bucket_params = {small.names[i] for i in range(small.buckets[small.fire_order[0]][2], small.buckets[small.fire_order[0]][3])}
    assert first_bucket_params <= {"b.weight", "b.bias"}            # the last layer's gradients are reduced first
    assert sorted(small.fire_order) == [0, 1, 2, 3]


def test_default_bucketing_and_optimizer_choice_on_the_host():
    """One process has nothing to overlap: the default is ONE bucket (one pack at the end of backward()); the 6 MB buckets are the
    default only with a process group of more than one rank (exercised by the gloo test below).  On CPU tensors the trainer keeps
    torch.optim.Adam; the fused clip + Adam kernels are a GPU-only path that refuses host buffers instead of falling back."""
    from domain_specific_image_compression_b200 import SicError
    m = Toy()
    tr = FlatTrainer(m, exclude=["gamma"])
    assert len(tr.buckets) == 1 and tr.buckets[0][:2] == (0, tr.flat.numel())
    assert not tr.fused and tr.opt is not None
    m2 = Toy()
    forced = FlatTrainer(m2, exclude=["gamma"], fused=True)
    with pytest.raises(SicError):
        forced.step(lambda: _loss(m2, torch.rand(2, 3, 8, 8)))


def test_dead_parameters_follow_the_gdn_mode():
    """Per GDN site: the reference's diagonal path trains `gamma_conv.weight` and never touches the CxC `gamma` (layers.py:13,21);
    GDN(dense=True) trains `gamma` and never touches `gamma_conv.weight`.  The bucket must hold exactly the live ones."""
    from domain_specific_image_compression_b200.layers import GDN
    from domain_specific_image_compression_b200.trainer import _dead_parameter_names
    m = torch.nn.Sequential(GDN(4), torch.nn.Sequential(GDN(4, inverse=True, dense=True)))
    assert _dead_parameter_names(m) == {"0.gamma", "1.0.gamma_conv.weight"}
    tr = FlatTrainer(m, fused=False)
    assert tr.names == ["0.beta", "0.gamma_conv.weight", "1.0.beta", "1.0.gamma"]


def test_missing_gradient_is_reported():
    m = Toy()
    tr = FlatTrainer(m, fused=False, exclude=["gamma"])
    with pytest.raises(RuntimeError):
        tr.step(lambda: m.a(torch.rand(1, 3, 4, 4)).sum())      # m.b gets no gradient


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                  # different initial weights per rank: the broadcast must fix that
    m = Toy()
    tr = FlatTrainer(m, lr=1e-2, grad_clip=0.5, fused=False, exclude=["gamma"])
    g = torch.Generator().manual_seed(7)
    x_all = torch.rand(4, 3, 8, 8, generator=g)
    x = x_all[rank * 2:(rank + 1) * 2]             # patches shard across ranks
    for _ in range(3):
        tr.step(lambda: _loss(m, x))
    q.put((rank, tr.flat.detach().numpy().copy()))      # by value: the worker may exit before the parent reads
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_equal_single_process_on_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = {k: torch.from_numpy(v) for k, v in got.items()}
    assert torch.equal(got[0], got[1])             # replicas stay identical
    torch.manual_seed(100)                         # rank 0's initial weights
    m = Toy()
    tr = FlatTrainer(m, lr=1e-2, grad_clip=0.5, fused=False, exclude=["gamma"])
    x_all = torch.rand(4, 3, 8, 8, generator=torch.Generator().manual_seed(7))
    for _ in range(3):
        tr.step(lambda: _loss(m, x_all))           # mean over 4 patches == mean of the two ranks' means over 2 patches
    assert torch.allclose(tr.flat.detach(), got[0], atol=2e-6)
