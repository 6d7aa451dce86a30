"""Multi-GPU host logic: band geometry, texture assignment, and a world_size-2 gloo run
on CPU in which each rank encodes its band (the oracle stands in for the GPU encoder),
the slices are written into one .astc file by offset and gathered, and the result must
be byte-identical to the single-rank encode."""
import os
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT


@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("size", [(1024, 1024), (250, 187), (37, 23), (8192, 8192), (5, 3), (1, 1)])
@pytest.mark.parametrize("parts", [1, 2, 3, 4, 8])
def test_band_plan_covers_exactly(native, dim, size, parts):
    from astc_encoder_b200 import sharding
    w, h = size
    opt = native.encode_option(is4x4=(dim == 4), is6x6=(dim == 6))
    plan = sharding.band_plan(w, h, opt, parts)
    bx, by = native.block_counts(w, h, opt)
    assert (bx, by) == ((w + dim - 1) // dim, (h + dim - 1) // dim)
    y, off = 0, 0
    for b in plan:
        assert b.y0 == y and b.byte_offset == off
        assert b.y0 % dim == 0 or b.rows == 0
        assert b.nbytes == ((b.rows + dim - 1) // dim) * bx * 16
        y += b.rows
        off += b.nbytes
    assert y == h and off == native.output_size(w, h, opt) == bx * by * 16
    # only the last non-empty band may hold a partial block row
    nonempty = [b for b in plan if b.rows]
    assert all(b.rows % dim == 0 for b in nonempty[:-1])
    # balanced to within one block row
    rows = [(b.rows + dim - 1) // dim for b in plan]
    assert max(rows) - min(rows) <= 1


def test_band_rejects_bad_arguments(native):
    import ctypes as C
    o = native.encode_option()._abi()
    L = native.lib()
    assert L.astc_b200_band(16, 16, C.byref(o), 0, 0, None, None, None, None) == -1
    assert L.astc_b200_band(16, 16, C.byref(o), 2, 2, None, None, None, None) == -1
    assert L.astc_b200_band(-1, 16, C.byref(o), 2, 0, None, None, None, None) == -1


def test_assign_textures_lpt():
    from astc_encoder_b200 import sharding
    # 512 mip chains of equal size deal out evenly
    chains = [5592405] * 512
    own = sharding.assign_textures(chains, 8)
    assert sorted(i for o in own for i in o) == list(range(512))
    assert {len(o) for o in own} == {64}
    # mixed sizes: LPT keeps the makespan within 4/3 of the ideal
    rng = np.random.default_rng(3)
    sizes = [int(s) for s in rng.integers(1, 1 << 22, 200)]
    own = sharding.assign_textures(sizes, 8)
    loads = [sum(sizes[i] for i in o) for o in own]
    assert sorted(i for o in own for i in o) == list(range(200))
    assert max(loads) <= (4 / 3) * (sum(sizes) / 8) + max(sizes)
    assert all(o == sorted(o) for o in own)


def _worker(rank, world, port, tmp, dim, w, h):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import astc_encoder_b200 as A
    from astc_encoder_b200 import sharding, synth
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        opt = A.encode_option(is4x4=(dim == 4), is6x6=(dim == 6), has_alpha=True)
        img = synth.synth_rgba(w, h, 77).numpy()
        enc = lambda rows, o: O.encode_image(rows, block_dim=dim, has_alpha=True)   # CPU stand-in for the GPU encoder
        b, blocks = sharding.encode_band(img, opt, rank, world, encode_fn=enc)
        path = os.path.join(tmp, "sharded.astc")
        if rank == 0:
            sharding.write_astc_sharded(path, w, h, opt, b, blocks, rank)
        dist.barrier()                                   # header + size exist before other ranks write
        if rank != 0:
            sharding.write_astc_sharded(path, w, h, opt, b, blocks, rank)
        dist.barrier()
        full = sharding.gather_blocks(blocks, w, h, opt, dst=0)
        if rank == 0:
            want = O.encode_image(img, block_dim=dim, has_alpha=True)
            assert np.array_equal(full, want)
            ref = os.path.join(tmp, "single.astc")
            A.save_astc(ref, dim, dim, w, h, want)
            assert open(ref, "rb").read() == open(path, "rb").read()
            open(os.path.join(tmp, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,w,h", [(4, 200, 96), (6, 250, 187), (4, 64, 4)])
def test_two_rank_gloo_bands_match_single(native, oracle, dim, w, h):
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 300) + dim
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, port, tmp, dim, w, h), nprocs=2, join=True)
        assert os.path.exists(os.path.join(tmp, "ok"))


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [4, 6])
@pytest.mark.parametrize("parts", [2, 3, 8])
def test_gpu_bands_concatenate_to_full_encode(native, dim, parts):
    """Each band through the C-ABI host entry point; concatenation == the one-shot device encode."""
    import torch
    from astc_encoder_b200 import sharding, synth
    w, h = 510, 383
    opt = native.encode_option(is4x4=(dim == 4), is6x6=(dim == 6), has_alpha=True, srgb=True)
    img = synth.synth_rgba(w, h, 99)
    full = native.read_gpu(native.encode_astc(img.cuda(), opt))
    got = np.concatenate([sharding.encode_band(img.numpy(), opt, r, parts)[1] for r in range(parts)])
    assert np.array_equal(got, full)


def test_cpulist_parser_and_numa_binding_never_raises():
    from astc_encoder_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding._parse_cpulist("") == set()
    info = sharding.bind_host_to_gpu(0)          # no GPU here: reports the error instead of raising
    assert isinstance(info, dict) and "bound" in info
