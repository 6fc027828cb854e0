"""Patch-parallel compress/decompress (rank r codes patches r, r+W, ...; host gather, no data-path collective).
CPU part: the sharding / merge logic over real `gloo` process groups with a stand-in codec (the product codec is CUDA only).
GPU part: the real model, shards merged in-process, round trip against forward()."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from domain_specific_image_compression_b200 import codec_parallel as CP


class FakeCodec:
    """compress(): per-patch byte strings derived from the patch content; decompress(): inverts them.  Same dict layout as
    CompressionModel.compress (custom_compress, eval_selfcontained_entropy.py:68-74)."""

    def compress(self, x, tail=10, coder="gpu"):
        B = x.size(0)
        q = (x * 255).round().to(torch.uint8)
        return {"strings": [[q[b, :1].numpy().tobytes(), q[b].numpy().tobytes()] for b in range(B)],
                "shape_y": [B, 3, x.size(2), x.size(3)], "shape_z": [B, 1, x.size(2), x.size(3)],
                "min_y": [int(q[b].min()) - tail for b in range(B)], "max_y": [int(q[b].max()) + tail for b in range(B)],
                "min_z": [int(q[b, :1].min()) - tail for b in range(B)], "max_z": [int(q[b, :1].max()) + tail for b in range(B)]}

    def decompress(self, compressed, coder="gpu"):
        B = len(compressed["strings"])
        shp = compressed["shape_y"][1:]
        out = torch.empty(B, *shp)
        for b in range(B):
            out[b] = torch.from_numpy(np.frombuffer(compressed["strings"][b][1], np.uint8).reshape(shp).copy()).float() / 255
        return out


def _batch(n):
    return torch.rand(n, 3, 4, 5, generator=torch.Generator().manual_seed(n))


def test_patch_indices_cover_every_patch_once():
    for n in (0, 1, 5, 8):
        for w in (1, 2, 3, 8):
            got = sorted(i for r in range(w) for i in CP.patch_indices(n, r, w))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        CP.patch_indices(4, 2, 2)


@pytest.mark.parametrize("n,world", [(5, 2), (2, 3), (6, 3), (1, 1)])
def test_merge_of_shards_equals_single_process(n, world):
    codec, x = FakeCodec(), _batch(n)
    whole = codec.compress(x)
    parts = [CP.compress_sharded(codec, x, gather=False, rank=r, world=world) for r in range(world)]
    merged = CP.merge_compressed(parts, n)
    assert merged == whole
    for r in range(world):
        idx = CP.patch_indices(n, r, world)
        assert CP.split_compressed(merged, idx) == (parts[r][1] if idx else {**{k: [] for k in CP._PER_PATCH_KEYS},
                                                                           "shape_y": [0, 3, 4, 5], "shape_z": [0, 1, 4, 5]})


def test_merge_detects_missing_and_duplicate_patches():
    codec, x = FakeCodec(), _batch(4)
    p0 = CP.compress_sharded(codec, x, gather=False, rank=0, world=2)
    with pytest.raises(ValueError, match="not produced"):
        CP.merge_compressed([p0], 4)
    with pytest.raises(ValueError, match="twice"):
        CP.merge_compressed([p0, p0], 4)
    with pytest.raises(ValueError, match="does not match"):
        CP.merge_compressed([(p0[0], None)], 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    codec, x = FakeCodec(), _batch(n)
    merged = CP.compress_sharded(codec, x)                       # rank/world from the process group, host gather
    x_hat = CP.decompress_sharded(codec, merged)
    q.put((rank, merged, x_hat.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(5, 2), (2, 3)])
def test_gloo_ranks_gather_the_single_process_result(n, world):
    """world_size 2 with a ragged split, world_size 3 with an idle rank: every rank ends with the single-process dict and
    the full reconstruction, in patch order."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    codec, x = FakeCodec(), _batch(n)
    whole = codec.compress(x)
    want = codec.decompress(whole).numpy()
    assert sorted(r for r, _, _ in got) == list(range(world))
    for _, merged, x_hat in got:
        assert merged == whole
        np.testing.assert_array_equal(x_hat, want)


@pytest.mark.gpu
def test_sharded_round_trip_on_the_real_model():
    """Two shards coded by the real CUDA path and merged in-process; each shard's reconstruction equals the forward pass of
    that shard bit for bit (the same property tests/test_gpu_model.py checks for a whole batch)."""
    import domain_specific_image_compression_b200 as sic
    torch.manual_seed(5)
    m = sic.CompressionModel(N=16, M=24, spatial_params=False, min_nu=2.0, max_nu=100.0).cuda().eval()
    with torch.no_grad():
        m.g_a.g_a[14].weight.mul_(30.0)
        m.h_a.h_a[6].weight.mul_(30.0)
    x = torch.rand(5, 3, 64, 64, device="cuda")
    world = 2
    parts = [CP.compress_sharded(m, x, gather=False, rank=r, world=world) for r in range(world)]
    merged = CP.merge_compressed(parts, 5)
    assert len(merged["strings"]) == 5 and merged["shape_y"][0] == 5
    for r in range(world):
        idx, x_hat = CP.decompress_sharded(m, merged, gather=False, rank=r, world=world)
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True):
            want = m(x[idx], quant_mode="round")["x_hat"].clamp(0, 1)
        assert torch.equal(x_hat, want)


# ------------------------------------------------------------------------------------------------ SIC-CONT-1 container
def test_container_round_trip_and_damage_detection():
    from domain_specific_image_compression_b200 import container as K
    codec, x = FakeCodec(), _batch(3)
    comp = codec.compress(x)
    blob = K.pack(comp)
    assert blob[:4] == b"SICC" and len(blob) == 36 + 3 * 24 + sum(len(z) + len(y) for z, y in comp["strings"]) + 4
    assert K.unpack(blob) == comp
    assert K.unpack(K.pack(CP.split_compressed(comp, []))) == CP.split_compressed(comp, [])      # zero patches
    for bad in (blob[:-1], blob[:50], b"XXXX" + blob[4:], blob[:40] + bytes([blob[40] ^ 1]) + blob[41:], blob + b"\0"):
        with pytest.raises(K.ContainerError):
            K.unpack(bad)
    with pytest.raises(K.ContainerError):
        K.pack({**comp, "min_y": comp["min_y"][:-1]})
    # negative supports survive (signed fields)
    neg = {**comp, "min_y": [-37, -1, 0], "min_z": [-3000, 5, -9]}
    assert K.unpack(K.pack(neg)) == neg
    # ... but a support no table can hold (wider than 4096 symbols, or max < min) is refused although the CRC is right
    for k, v in (("min_z", [-2 ** 31, 5, -9]), ("max_y", [-40, 264, 264])):
        with pytest.raises(K.ContainerError):
            K.unpack(K.pack({**comp, k: v}))


def test_gather_order_inverts_the_round_robin_sharding():
    """decompress_sharded over NCCL gathers equal-sized, zero-padded blocks (rank r: patches r, r+W, ...) and re-orders them with
    gather_order: for every batch size and world size, including ragged and idle ranks, that re-ordering is the identity on patch ids."""
    from domain_specific_image_compression_b200 import codec_parallel as CP
    for world in (1, 2, 3, 4, 8):
        for B in (1, 2, 5, 8, 13, 16, 17):
            per = (B + world - 1) // world
            blocks = []
            for r in range(world):
                idx = CP.patch_indices(B, r, world)
                blocks += idx + [-1] * (per - len(idx))            # -1 = padding rows of the all_gather block
            got = [blocks[k] for k in CP.gather_order(B, world)]
            assert got == list(range(B)), (world, B)


def test_training_only_fast_paths_are_off_by_default_and_gated():
    """The library default must reproduce the reference bit for bit: the three training-only switches are off, and the layer
    predicates refuse eval mode, CPU tensors and shapes the kernels do not cover."""
    import torch
    import torch.nn as nn
    from domain_specific_image_compression_b200 import layers as L, model as M
    assert L.FUSE_FIRST_LAYER is False and L.FAST_LAST_LAYER is False and M.OVERLAP_HYPER_BRANCH is False
    conv, gdn = nn.Conv2d(3, 128, 3, 1, 1), L.GDN(128)
    x = torch.rand(1, 3, 16, 16)
    try:
        L.FUSE_FIRST_LAYER = L.FAST_LAST_LAYER = True
        assert not L._first_layer_fusable(conv, gdn, x, True)                       # CPU tensor
        assert not L._first_layer_fusable(conv, gdn, x, False)                      # eval mode
        assert not L._last_layer_fast(nn.ConvTranspose2d(128, 3, 5, 2, 2, output_padding=1), x, True)    # CPU tensor
        assert not L._last_layer_fast(nn.ConvTranspose2d(128, 3, 5, 2, 2, output_padding=1), x, False)
    finally:
        L.FUSE_FIRST_LAYER = L.FAST_LAST_LAYER = False
