"""(e) tail of the training step on the GPU: the two-launch clip + Adam of csrc/train_step.cu against the reference's step
(train.py:182-183 optim.Adam, :200-203 clip_grad_norm_ + step) - against a float64 statement of the update rule element by element,
and through FlatTrainer against torch.optim.Adam + clip_grad_norm_ on an unflattened copy of the same module."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _adam_f64(p, g, m, v, t, inv_world, clip, lr, b1, b2, eps, wd):
    # the hyper-parameters as the kernel (and torch's fused Adam) sees them: rounded to float32 (1 - float32(0.999) != 0.001)
    inv_world, clip, lr, b1, b2, eps, wd = (float(np.float32(h)) for h in (inv_world, clip, lr, b1, b2, eps, wd))
    g = g * inv_world
    norm = float(np.sqrt((g * g).sum()))
    coef = min(1.0, clip / (norm + 1e-6)) if clip > 0 else 1.0
    g = g * coef
    if wd:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    p = p - (lr / (1 - b1 ** t)) * m / (np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps)
    return p, m, v, norm


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 65537, 6_400_003])
@pytest.mark.parametrize("inv_world,clip,wd", [(1.0, 1.0, 0.0), (0.125, 0.5, 0.0), (1.0, 0.0, 1e-2), (0.5, 1e6, 0.0)])
def test_clip_adam_step_vs_float64(n, inv_world, clip, wd):
    from domain_specific_image_compression_b200 import functional as F
    gen = torch.Generator(device="cuda").manual_seed(n)
    p = torch.randn(n, device="cuda", generator=gen) * 1e-3       # small parameters: fp32 rounding of p itself stays far below one step
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    step = torch.zeros((), device="cuda")
    norm = torch.zeros((), device="cuda")
    ws = F.clip_adam_workspace(n, p.device)
    p64, m64, v64 = p.double().cpu().numpy(), m.double().cpu().numpy(), v.double().cpu().numpy()
    lr, betas, eps = 1e-3, (0.9, 0.999), 1e-8
    for t in range(1, 4):
        g = torch.randn(n, device="cuda", generator=gen) * (3.0 / max(n, 1) ** 0.5) / inv_world
        g_before = g.clone()
        F.clip_adam_step(p, g, m, v, step, norm, ws, inv_world=inv_world, clip=clip, lr=lr, betas=betas, eps=eps, weight_decay=wd)
        p64, m64, v64, n64 = _adam_f64(p64, g.double().cpu().numpy(), m64, v64, t, inv_world, clip, lr, *betas, eps, wd)
        assert torch.equal(g, g_before)                                      # the gradient buffer is read only
        assert float(step) == t
        assert abs(float(norm) - n64) <= 2e-6 * n64 + 1e-12
        # one Adam step moves a parameter by <= ~lr: compare the MOVE, relative to lr, not the parameter
        assert float(np.abs(p.double().cpu().numpy() - p64).max()) <= 2e-5 * lr
        assert float(np.abs(m.double().cpu().numpy() - m64).max()) <= 1e-6 * float(np.abs(m64).max()) + 1e-12
        assert float(np.abs(v.double().cpu().numpy() - v64).max()) <= 1e-6 * float(np.abs(v64).max()) + 1e-12


def test_clip_adam_step_rejects_bad_arguments():
    from domain_specific_image_compression_b200 import _lib
    from domain_specific_image_compression_b200 import functional as F
    p = torch.zeros(16, device="cuda")
    step, ws = torch.zeros((), device="cuda"), F.clip_adam_workspace(16, p.device)
    kw = dict(inv_world=1.0, clip=1.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    with pytest.raises(_lib.SicError):
        F.clip_adam_step(p, p[:8], p.clone(), p.clone(), step, None, ws, **kw)                  # size mismatch
    with pytest.raises(_lib.SicError):
        F.clip_adam_step(p.cpu(), p.cpu(), p.cpu(), p.cpu(), step, None, ws, **kw)              # no CPU path
    with pytest.raises(_lib.SicError):
        F.clip_adam_step(p, p.clone(), p.clone(), p.clone(), step, None, ws, **{**kw, "betas": (1.0, 0.999)})


def test_pack_flat_equals_cat():
    """The gradient pack: 300 tensors of odd and large sizes (slices off the 16-byte grid, several launches of 128, chunks of 4096
    floats with ragged ends) land where torch.cat puts them; untouched gaps stay untouched."""
    from domain_specific_image_compression_b200 import _lib
    from domain_specific_image_compression_b200 import functional as F
    gen = torch.Generator(device="cuda").manual_seed(9)
    sizes = [1, 3, 4, 5, 4096, 4097, 8191, 100_003, 1_000_000] + [int(v) for v in torch.randint(1, 20_000, (291,), generator=torch.Generator().manual_seed(1))]
    ts = [torch.randn(n, device="cuda", generator=gen) for n in sizes]
    offs, off = [], 7                                            # start off the grid on purpose, leave a gap of 2 between tensors
    for n in sizes:
        offs.append(off)
        off += n + 2
    dst = torch.full((off + 5,), -7.0, device="cuda")
    F.pack_flat(ts, offs, dst)
    ref = torch.full_like(dst, -7.0)
    for t, o in zip(ts, offs):
        ref[o:o + t.numel()] = t
    assert torch.equal(dst, ref)
    with pytest.raises(_lib.SicError):
        F.pack_flat([ts[0]], [dst.numel()], dst)                 # slice outside the destination
    with pytest.raises(_lib.SicError):
        F.pack_flat([ts[0].cpu()], [0], dst)


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.b = torch.nn.Conv2d(8, 5, 3, padding=1)
        self.gamma = torch.nn.Parameter(torch.eye(4))       # dead parameter, like GDN's CxC gamma

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def test_flat_trainer_fused_tail_equals_reference_step():
    """FlatTrainer on the GPU (default: the fused tail) vs zero_grad / backward / clip_grad_norm_ / torch.optim.Adam on a copy, eager
    and as a captured CUDA graph (the update counter and the norm live on the device: replays keep counting)."""
    from domain_specific_image_compression_b200.trainer import FlatTrainer
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(0)
        m = Toy().cuda()
        ref = copy.deepcopy(m)
        tr = FlatTrainer(m, lr=1e-2, grad_clip=1.0, exclude=["gamma"])
        assert tr.fused and tr.opt is None
        live = [p for n, p in ref.named_parameters() if n != "gamma"]
        opt = torch.optim.Adam(live, lr=1e-2)
        x = torch.rand(4, 3, 8, 8, device="cuda")
        loss_fn = lambda mod: (mod(x) - 0.3).square().mean() * 50.0

        def ref_step():
            opt.zero_grad(set_to_none=True)
            loss_fn(ref).backward()
            n = torch.nn.utils.clip_grad_norm_(live, 1.0)
            opt.step()
            return float(n)

        for _ in range(3):
            tr.step(lambda: loss_fn(m))
            n_ref = ref_step()
            assert abs(float(tr.grad_norm) - n_ref) <= 1e-5 * n_ref
        replay = tr.capture(lambda: loss_fn(m), warmup=2)          # 2 eager warm-up steps (recording the graph executes nothing)
        for _ in range(2):
            ref_step()
        for _ in range(2):
            replay()
            ref_step()
        torch.cuda.synchronize()
        assert float(tr.step_count) == 3 + 2 + 2
        for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            assert float((p - q).abs().max()) <= 2e-5, n
        tr.release_graph()
    finally:
        torch.backends.cudnn.allow_tf32 = old
