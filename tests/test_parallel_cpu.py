"""Host-side multi-GPU logic on CPU: tile ownership, slab layout, and a world_size-2 gloo run in which each rank
renders only its tiles (with the oracle standing in for the device) and the gathered frame equals the single-rank
frame bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

from metal4_raytracing_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("w,h,n", [(64, 64, 2), (100, 70, 4), (1920, 1080, 8), (33, 17, 3)])
def test_tile_partition_is_a_partition(w, h, n):
    masks = [parallel.owner_mask(w, h, n, r) for r in range(n)]
    total = np.sum(masks, axis=0)
    assert np.all(total == 1)
    tx, ty = parallel.tile_grid(w, h)
    owned = [parallel.owned_tiles(w, h, n, r) for r in range(n)]
    assert sorted(sum(owned, [])) == list(range(tx * ty))
    assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    assert parallel.slab_tiles(w, h, n) == max(len(o) for o in owned)


@pytest.mark.parametrize("w,h,n", [(64, 48, 2), (100, 70, 4), (33, 17, 3)])
def test_pack_unpack_roundtrip(w, h, n):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 65535, (h, w, 4)).astype(np.uint16)
    slabs = np.stack([parallel.pack_tiles_host(img * parallel.owner_mask(w, h, n, r)[..., None].astype(np.uint16), n, r)
                      for r in range(n)])
    assert np.array_equal(parallel.unpack_tiles_host(slabs, w, h, n), img)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from metal4_raytracing_b200 import scene
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 80, 56
    sc, u, seed = scene.Scene.named("K3small", w, h, assets=None)
    u.samplesPerPixel, u.maxBounces = 1, 2
    seeds = scene.seed_image(w, h, seed)
    orc = oracle.Oracle(sc, threads=2)
    imgs = oracle.FrameImages(w, h, seeds)
    frames = []
    for f in range(2):  # two frames: history stays on the owning rank
        u.frameIndex = f
        orc.render(u, imgs, tile_modulo=world, tile_remainder=rank)
        frames.append(parallel.gather_frame_host(imgs.output, world, rank, dist))
        imgs.swap()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack(frames))
    dist.destroy_process_group()


def test_two_rank_gloo_frame_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    import oracle
    from metal4_raytracing_b200 import scene
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w, h = 80, 56
    sc, u, seed = scene.Scene.named("K3small", w, h, assets=None)
    u.samplesPerPixel, u.maxBounces = 1, 2
    imgs = oracle.FrameImages(w, h, scene.seed_image(w, h, seed))
    orc = oracle.Oracle(sc, threads=2)
    ref = []
    for f in range(2):
        u.frameIndex = f
        orc.render(u, imgs)
        ref.append(imgs.output.copy())
        imgs.swap()
    ref = np.stack(ref)
    for r in range(2):
        got = np.load(os.path.join(tmp_path, f"rank{r}.npy"))
        assert np.array_equal(got.view(np.uint16), ref.view(np.uint16)), f"rank {r}"
