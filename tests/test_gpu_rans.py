"""N1: the GPU rANS coder (one warp per stream) against the host coder and the C oracle: identical bytes, round trips,
erasure detection; compress()/decompress() with either coder."""
import numpy as np
import pytest
import torch

from oracle import clib

pytestmark = pytest.mark.gpu


def _F():
    from domain_specific_image_compression_b200 import functional as F
    return F


def _case(rng, S, C, hw, Ls):
    stride = max(Ls) + 1
    tabs, syms = [], []
    for s in range(S):
        L = Ls[s]
        sig = np.exp(rng.normal(0, 1, C)).astype(np.float32)
        nu = (2 + rng.random(C) * 20).astype(np.float32)
        mn = np.array([-(L // 2)], np.int32)
        t = clib.build_tables("studentt", sig, nu, np.zeros(C, np.int32), mn, mn + L - 1)
        full = np.zeros((C, stride), np.uint16)
        full[:, :L + 1] = t
        tabs.append(full)
        syms.append(np.clip(np.rint(rng.standard_t(3, (C, hw)) * sig[:, None]).astype(np.int64) - mn[0], 0, L - 1).astype(np.int32))
    return np.stack(tabs), np.stack(syms)


@pytest.mark.parametrize("S,C,hw,Ls", [(1, 1, 1, [2]), (3, 4, 32, [21, 40, 9]), (2, 7, 33, [31, 31]), (5, 24, 64, [25, 60, 33, 21, 47]),
                                        (2, 192, 1024, [27, 35])])
def test_device_coder_bytes_equal_host_and_oracle(S, C, hw, Ls):
    F = _F()
    rng = np.random.default_rng(S * 100 + C)
    tabs, syms = _case(rng, S, C, hw, Ls)
    n = C * hw
    out, nbytes = F.rans_encode_device(torch.from_numpy(syms.reshape(S, n)).cuda(), torch.from_numpy(tabs.reshape(S * C, -1)).cuda(),
                                       torch.tensor(Ls, dtype=torch.int32).cuda(), hw, C)
    # the single-kernel encoder (sic_rans_encode, no workspace) and the two-phase one (sic_rans_encode_ws, what rans_encode_device runs)
    # must emit the same bytes
    import ctypes
    from domain_specific_image_compression_b200 import _lib
    lib = _lib.load()
    sym_d, tab_d = torch.from_numpy(syms.reshape(S, n)).cuda(), torch.from_numpy(tabs.reshape(S * C, -1)).cuda()
    Ls_d = torch.tensor(Ls, dtype=torch.int32).cuda()
    out1, nb1 = torch.zeros_like(out), torch.zeros_like(nbytes)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.sic_rans_encode(vp(sym_d), vp(tab_d), vp(Ls_d), S, n, hw, C, tab_d.shape[-1], vp(out1), out1.shape[1], vp(nb1),
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0, lib.sic_last_error()
    assert torch.equal(nb1, nbytes)
    for s in range(S):
        assert torch.equal(out1[s, :int(nb1[s])], out[s, :int(nbytes[s])])
    out, nbytes = out.cpu().numpy(), nbytes.cpu().numpy()
    for s in range(S):
        ref = clib.rans_encode(syms[s], tabs[s][:, :Ls[s] + 1], Ls[s], hw)
        assert F.rans_encode(syms[s], tabs[s], Ls[s], hw) == ref
        assert out[s, :nbytes[s]].tobytes() == ref, f"stream {s}"
    sym, status = F.rans_decode_device(torch.from_numpy(out).cuda(), torch.from_numpy(nbytes).cuda(), torch.from_numpy(tabs.reshape(S * C, -1)).cuda(),
                                       torch.tensor(Ls, dtype=torch.int32).cuda(), n, hw, C)
    assert status.abs().max().item() == 0 and np.array_equal(sym.cpu().numpy(), syms.reshape(S, n))


def test_device_coder_spatial_rows_and_errors():
    """sym_per_row = 1 (one table per symbol: spatial_params=True), truncated streams, out-of-range symbols."""
    F = _F()
    rng = np.random.default_rng(3)
    n, L = 200, 17
    sig = np.exp(rng.normal(0, 1, n)).astype(np.float32)
    nu = (2 + rng.random(n) * 20).astype(np.float32)
    mn = np.array([-8], np.int32)
    tab = clib.build_tables("studentt", sig, nu, np.zeros(n, np.int32), mn, mn + L - 1)
    sym = np.clip(np.rint(rng.standard_t(3, n) * sig).astype(np.int64) + 8, 0, L - 1).astype(np.int32)
    out, nb = F.rans_encode_device(torch.from_numpy(sym[None]).cuda(), torch.from_numpy(tab).cuda(), torch.tensor([L], dtype=torch.int32).cuda(), 1, n)
    ref = clib.rans_encode(sym, tab, L, 1)
    assert out[0, :int(nb[0])].cpu().numpy().tobytes() == ref
    dec, st = F.rans_decode_device(out, nb, torch.from_numpy(tab).cuda(), torch.tensor([L], dtype=torch.int32).cuda(), n, 1, n)
    assert int(st[0]) == 0 and np.array_equal(dec[0].cpu().numpy(), sym)
    big = np.tile(sym, 40)
    tabb = np.tile(tab, (40, 1))
    out, nb = F.rans_encode_device(torch.from_numpy(big[None]).cuda(), torch.from_numpy(tabb).cuda(), torch.tensor([L], dtype=torch.int32).cuda(), 1, big.size)
    assert int(nb[0]) > 200
    _, st = F.rans_decode_device(out, nb - 4, torch.from_numpy(tabb).cuda(), torch.tensor([L], dtype=torch.int32).cuda(), big.size, 1, big.size)
    assert int(st[0]) == -5                                            # SIC_E_TRUNCATED
    _, st = F.rans_decode_device(out, torch.tensor([100], dtype=torch.int32).cuda(), torch.from_numpy(tabb).cuda(), torch.tensor([L], dtype=torch.int32).cuda(), big.size, 1, big.size)
    assert int(st[0]) == -5
    bad = sym.copy(); bad[5] = L
    _, nb = F.rans_encode_device(torch.from_numpy(bad[None]).cuda(), torch.from_numpy(tab).cuda(), torch.tensor([L], dtype=torch.int32).cuda(), 1, n)
    assert int(nb[0]) == -1


def test_device_decoder_reports_damage_and_bad_supports():
    """Flipped bits / stray words must come back as SIC_E_CORRUPT (-6) or SIC_E_TRUNCATED (-5), never as symbols; a support
    L <= 0 (min > max in a hand-made container) or L wider than the table row is SIC_E_BADARG (-1) and touches no memory."""
    F = _F()
    rng = np.random.default_rng(5)
    S, C, hw, L = 4, 6, 64, 9
    n = C * hw
    pm = rng.random((S, C, L)) + 0.05
    cdf = np.concatenate([np.zeros((S, C, 1)), np.cumsum(pm / pm.sum(-1, keepdims=True), -1)], -1)
    cdf[..., -1] = 1.0
    tabs = np.zeros((S, C, L + 3), np.uint16)
    tabs[..., :L + 1] = (cdf * 65535).astype(np.uint16)
    syms = rng.integers(0, L, (S, n)).astype(np.int32)
    t_d = torch.from_numpy(tabs.reshape(S * C, -1)).cuda()
    Ls = torch.full((S,), L, dtype=torch.int32).cuda()
    out, nb = F.rans_encode_device(torch.from_numpy(syms).cuda(), t_d, Ls, hw, C)
    dec, st = F.rans_decode_device(out, nb, t_d, Ls, n, hw, C)
    assert st.abs().max().item() == 0 and np.array_equal(dec.cpu().numpy(), syms)
    hurt = out.clone()
    hurt[0, 5] ^= 0x40                                   # a state word
    hurt[1, 140] ^= 0x01                                 # an early renormalisation word
    hurt[2, int(nb[2]) - 1] ^= 0x80                      # the last word
    nb2 = nb.clone(); nb2[3] += 2                        # one stray (zero) word after stream 3
    _, st = F.rans_decode_device(hurt, nb2, t_d, Ls, n, hw, C)
    st = st.cpu().numpy()
    assert all(int(v) in (-5, -6) for v in st), st
    badL = torch.tensor([0, -3, L + 3, 5000], dtype=torch.int32).cuda()
    _, st = F.rans_decode_device(out, nb, t_d, badL, n, hw, C)
    assert st.cpu().tolist() == [-1, -1, -1, -1]
    _, nbe = F.rans_encode_device(torch.from_numpy(syms).cuda(), t_d, badL, hw, C)
    assert nbe.cpu().tolist() == [-1, -1, -1, -1]
    torch.cuda.synchronize()


def test_compress_gpu_and_host_coders_agree(golden):
    import domain_specific_image_compression_b200 as sic
    G = golden("model_small")
    m = sic.CompressionModel(N=16, M=24, min_nu=2.0).cuda()
    m.load_state_dict({k[3:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("sd.")})
    m.eval()
    x = torch.nn.functional.interpolate(torch.rand(4, 3, 16, 16, generator=torch.Generator().manual_seed(2)), size=(128, 128), mode="bilinear").clamp(0, 1).cuda()
    a, b = m.compress(x, coder="gpu"), m.compress(x, coder="host")
    assert a == b                                                        # identical dicts: bytes, shapes, supports
    xa, xb = m.decompress(a, coder="gpu"), m.decompress(a, coder="host")
    assert torch.equal(xa, xb)
    a["strings"][2][1] = a["strings"][2][1][:-8]
    with pytest.raises(sic.SicError):
        m.decompress(a, coder="gpu")
