"""Multi-GPU frame assembly on real devices (needs >= 2 GPUs; skipped otherwise): each rank renders only its tiles
on its own B200 and the frame every rank ends up with — through NVLink peer stores from the resolve kernel, or
through rt_pack_tiles + NCCL all-gather + rt_unpack_tiles — equals the single-GPU frame bit for bit, over several
EMA frames of an animated scene."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, FRAMES = 200, 136, 3


def _scene(name):
    from metal4_raytracing_b200 import scene
    sc, u, seed = scene.Scene.named(name, W, H, assets=None)
    u.samplesPerPixel, u.maxBounces = 2, 3
    return sc, u, scene.seed_image(W, H, seed)


def _render(ctx, world, rank, mode, dist=None, name="K5small"):
    from metal4_raytracing_b200 import _abi as A, device, parallel
    sc, u, seeds = _scene(name)
    if mode == "samples":  # shares are summed in fp32; the motion-adaptive features depend on sample 0's owner
        u.samplesPerPixel = 4
        u.enableMotionAdaptiveSampling = u.enableMotionAdaptiveAccumulation = 0
    rnd = device.Renderer(ctx, sc, W, H, seeds=seeds, fp32=(mode == "samples"))
    xchg = parallel.FrameExchange(rnd, world, rank, mode=mode)
    frames = []
    for f in range(FRAMES):
        u.frameIndex = f
        if f:
            sc.animate(f / 60.0)
            rnd.update()
        rnd.draw(u, **xchg.draw_partition())
        xchg.finish_frame()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()
        frames.append(rnd.read_image(A.TEXTURE_ACCUMULATION).copy())
    xchg.close()
    rnd.close()
    return np.stack(frames)


def _worker(rank, world, port, mode, out_dir, name, own_stream):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from metal4_raytracing_b200 import device
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    ctx = device.Context(rank)
    if not own_stream:
        # the caller moves the library onto torch's stream (what bench.py does); with own_stream the library keeps
        # the stream rt_create gave it and FrameExchange has to order NCCL against that one (parallel.py)
        stream = torch.cuda.Stream(device=rank)
        torch.cuda.set_stream(stream)
        ctx.set_stream(stream.cuda_stream)
    np.save(os.path.join(out_dir, f"{mode}_rank{rank}.npy"), _render(ctx, world, rank, mode, dist, name))
    ctx.close()
    dist.destroy_process_group()


def _world():
    """Ranks to test with: every visible GPU up to RT_TEST_WORLD (default 2; the 8-rank run sets it to 8)."""
    import torch
    return min(torch.cuda.device_count(), int(os.environ.get("RT_TEST_WORLD", "2")))


@pytest.mark.parametrize("own_stream", [False, True], ids=["torch-stream", "library-stream"])
@pytest.mark.parametrize("name", ["K5small", "K3small"])
@pytest.mark.parametrize("mode", ["peer", "gather", "samples"])
def test_multi_gpu_frame_equals_single_gpu(tmp_path, mode, name, own_stream):
    world = _world()
    if world < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from metal4_raytracing_b200 import device
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, mode, str(tmp_path), name, own_stream), nprocs=world, join=True)
    ctx = device.Context(0)
    ref = _render(ctx, 1, 0, mode, name=name)
    ctx.close()
    for r in range(world):
        got = np.load(os.path.join(tmp_path, f"{mode}_rank{r}.npy"))
        if mode == "samples":  # the all-reduce reassociates float sums: equal to rounding, and identical on every rank
            assert np.abs(got.astype(np.float64) - ref).max() <= 4e-6 * max(1.0, float(np.abs(ref).max())), f"rank {r}"
            assert np.array_equal(got, np.load(os.path.join(tmp_path, f"{mode}_rank0.npy")))
        else:
            assert np.array_equal(got.view(np.uint16), ref.view(np.uint16)), f"{mode}/{name}: rank {r} of {world} differs"
