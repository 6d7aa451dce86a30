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
            dist.barrier()                               # worst case on purpose: rank 0 (header) writes LAST
        sharding.write_astc_sharded(path, w, h, opt, b, blocks, rank)
        if rank != 0:
            dist.barrier()
        dist.barrier()                                   # readers wait for the last writer
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


def _gpu_count() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("dim,w,h", [(4, 4096, 2048), (6, 2050, 1027)])
def test_gpu_bands_on_separate_devices_match_single_gpu_encode(native, dim, w, h):
    """North_star's band sharding with REAL encoders on >= 2 GPUs: band g lives on device g, is encoded there
    into that device's slice, and the slices are compared byte for byte ON the GPUs with device 0's encode
    of the whole texture (blocks are independent: ASTC_Encode.hlsl:561-581)."""
    import torch
    from astc_encoder_b200 import sharding, synth
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 CUDA devices (run under gpurun --gpus 2)")
    opt = native.encode_option(is4x4=(dim == 4), is6x6=(dim == 6), has_alpha=True)
    full_src = synth.synth_rgba(w, h, 4242, device="cuda:0")
    full = native.encode_astc(full_src, opt)
    plan = sharding.band_plan(w, h, opt, n)
    outs = []
    for g, b in enumerate(plan):
        slab = synth.synth_rgba(w, h, 4242, device=f"cuda:{g}", row0=b.y0, rows=b.rows)   # generated where it is encoded
        outs.append(native.encode_astc(slab, opt))
        assert outs[-1].device.index == g and outs[-1].numel() == b.nbytes
    for g in range(n):
        torch.cuda.synchronize(g)
    for b, o in zip(plan, outs):
        lo = b.byte_offset // 16
        assert torch.equal(full[lo:lo + b.nbytes // 16], o.to("cuda:0")), f"band {b.part}"


def _nccl_worker(rank, world, port, tmp, dim, w, h):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import astc_encoder_b200 as A
    from astc_encoder_b200 import sharding, synth
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        opt = A.encode_option(is4x4=(dim == 4), is6x6=(dim == 6), has_alpha=True, srgb=True)
        b = sharding.band_plan(w, h, opt, world)[rank]
        slab = synth.synth_rgba(w, h, 77, device="cuda", row0=b.y0, rows=b.rows)
        blocks = A.read_gpu(A.encode_astc(slab, opt))                      # this rank's band, encoded on its own GPU
        path = os.path.join(tmp, "sharded.astc")
        sharding.write_astc_sharded(path, w, h, opt, b, blocks, rank)      # any order, no barrier between writers
        dist.barrier()
        full = sharding.gather_blocks(blocks, w, h, opt, dst=0)            # optional convenience, over NCCL here
        if rank == 0:
            want = A.read_gpu(A.encode_astc(synth.synth_rgba(w, h, 77, device="cuda"), opt))
            assert np.array_equal(full, want)
            ref = os.path.join(tmp, "single.astc")
            A.save_astc(ref, dim, dim, w, h, want)
            assert open(ref, "rb").read() == open(path, "rb").read()
            open(os.path.join(tmp, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("dim,w,h", [(4, 2048, 1024), (6, 1030, 515)])
def test_two_rank_nccl_bands_on_gpus_match_single(native, dim, w, h):
    """One process per GPU (NCCL rendezvous), CUDA encoders, one shared .astc file written by offset."""
    if _gpu_count() < 2:
        pytest.skip("needs >= 2 CUDA devices (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = 29900 + (os.getpid() % 300) + dim
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_nccl_worker, args=(2, port, tmp, dim, w, h), nprocs=2, join=True)
        assert os.path.exists(os.path.join(tmp, "ok"))


def test_sharded_writer_is_order_independent(native, tmp_path):
    """astc_b200_save_astc_slice never truncates: the header writer may come last, slices in any order,
    and a stale longer file is cut to size."""
    from astc_encoder_b200 import sharding
    w, h, dim = 64, 40, 4
    opt = native.encode_option(has_alpha=True)
    rng = np.random.default_rng(5)
    blocks = rng.integers(0, 256, (native.output_size(w, h, opt) // 16, 16), dtype=np.uint8)
    plan = sharding.band_plan(w, h, opt, 3)
    path = tmp_path / "s.astc"
    path.write_bytes(b"x" * 100000)                                          # stale, longer than the result
    for rank in (2, 1, 0):
        b = plan[rank]
        sharding.write_astc_sharded(str(path), w, h, opt, b, blocks[b.byte_offset // 16:(b.byte_offset + b.nbytes) // 16], rank)
    ref = tmp_path / "ref.astc"
    native.save_astc(str(ref), dim, dim, w, h, blocks)
    assert path.read_bytes() == ref.read_bytes()


def test_cpulist_parser_and_numa_binding_never_raises():
    from astc_encoder_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding._parse_cpulist("") == set()
    info = sharding.bind_host_to_gpu(0)          # no GPU here: reports the error instead of raising
    assert isinstance(info, dict) and "bound" in info
