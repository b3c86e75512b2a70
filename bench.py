#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the ray-tracing hot path on B200 (BASELINE.json's metric).

A step is one frame of the workload: one dispatch of the path-tracing kernel over the whole render target
(rt_trace through the C-ABI), plus, for animated scenes, the skinning / BLAS refit / TLAS update that precedes it.
Workload (config.workload): BASELINE.json configs[2] — dragon stand-in (871,200 triangles) + two planes, 1920x1080,
16 spp, maxBounces 3, EMA accumulation over frames, default area + spot lights, environment lookup OFF (the
reference has none; the environment-lit variant is reported under "others"). The scene (67 MB of BVH
+ geometry) fits L2 only partly and every frame uses a new Halton index, so timed frames are not repeats of cached
work; an L2 flush between frames is also done.

  value     whole-job Mrays/s, device-timed, inputs resident in HBM (max over ranks for N > 1)
  e2e       the same metric through the public host API with HOST inputs: per frame rtr_update (pinned H2D of
            instance descriptors + lights, TLAS update) + draw + D2H of the finished frame, wall clock
  roofline  dominant kernel (k_wf_traverse, the persistent software traversal) vs the measured HBM peak: algorithmic
            bytes per SURVEY.md §8(d) (rays of the timed frames x bytes/ray) over that kernel's own launch time,
            measured live with CUDA events the library records around each of its launches during the timed region;
            `counted` restates it with the node / triangle / instance fetches a counter build of the kernel counted
            (profiles/work_counts.json), `issue` is the fraction of the machine's thread-instruction slots the kernel
            used in the committed ncu capture (profiles/issue.json) — the resource that actually binds it
  cpu_baseline  the CPU oracle (oracle/, a port of the reference kernels) on a bounded tile sample of the same frame
  others    short passes of the other BASELINE configurations in the same process (N = 1 only), outside the timed region
  frame_equal  N > 1: the frame rank 0 assembled from all ranks equals, bit for bit, the frame one rank renders alone

`--impl reference` times the oracle alone (the reference itself is Swift/Metal and cannot run on Linux).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene, width, height, spp, maxBounces)
    "K3": ("K3", 1920, 1080, 16, 3),
    "K3headline": ("K3", 1920, 1080, 1, 2),   # north_star's target configuration: 1 spp, primary + shadow + 1 bounce
    "K3glass": ("K3glass", 1920, 1080, 16, 3),  # the reference's own dragon material (Model.swift:22-26)
    "K3env": ("K3", 1920, 1080, 16, 3),       # + HDR environment (procedural 4096x2048 sky), sampled as a light
    "K2": ("K2", 1920, 1080, 4, 2),
    "K4": ("K4", 3840, 2160, 8, 2),
    "K5": ("K5", 1920, 1080, 2, 2),
    "K3small": ("K3small", 512, 512, 2, 3),
}
OTHERS = ["K3headline", "K3glass", "K3env", "K2", "K4", "K5"]
B_RAY = {"K3": 672.0, "K3headline": 672.0, "K3glass": 672.0, "K3env": 672.0, "K2": 592.0, "K4": 904.0, "K5": 592.0,
         "K3small": 512.0}
B_HIT, B_PIXEL, B_VERTEX = 300.0, 32.0, 120.0
B_NODE, B_TRIANGLE, B_INSTANCE = 80.0, 48.0, 64.0  # what one node step / triangle test / instance entry fetches


def _profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    return {}


def load_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    t = _profile_json("traffic.json").get(workload) or _profile_json("traffic.json").get(workload.replace("headline", ""))
    if t:
        return int(t["dram_bytes_per_launch"]), t["source"]
    return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            if self._stop.is_set():
                break
            parts = [p.strip() for p in line.split(",")]
            try:
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def build_scene(workload, assets="auto"):
    from metal4_raytracing_b200 import scene
    name, w, h, spp, mb = WORKLOADS[workload]
    sc, u, seed = scene.Scene.named(name, w, h, assets=assets)
    u.samplesPerPixel, u.maxBounces = spp, mb
    seeds = scene.seed_image(w, h, seed)
    return sc, u, seeds, w, h


def workload_text(workload):
    name, w, h, spp, mb = WORKLOADS[workload]
    env = ("procedural 4096x2048 HDR sky bound: lookup on a miss + sampled as a light with MIS"
           if workload == "K3env" else "env lookup off")
    return (f"{workload}: {name} scene, {w}x{h}, {spp} spp, maxBounces {mb}, EMA accumulation, {env}; "
            "dragon/bunny/robot are procedural stand-ins (assets absent from the mount)")


def run_oracle_sample(workload, steps, warmup, seconds_budget=20.0):
    """Times the CPU oracle on a bounded sample (a tile subset) of the workload's frame. Returns (line dict)."""
    import oracle
    sc, u, seeds, w, h = build_scene(workload)
    t0 = time.time()
    orc = oracle.Oracle(sc)
    build_s = time.time() - t0
    imgs = oracle.FrameImages(w, h, seeds)
    # calibrate: 1/64 of the tiles
    u.frameIndex = 0
    t0 = time.time()
    st, _ = orc.render(u, imgs, tile_modulo=64, tile_remainder=0)
    dt = max(1e-6, time.time() - t0)
    est_full = dt * 64
    per_step_budget = max(1.0, seconds_budget / max(1, steps + warmup))
    modulo = int(max(1, min(64, np.ceil(est_full / per_step_budget))))
    times, rays = [], []
    for i in range(warmup + steps):
        u.frameIndex = i
        t0 = time.time()
        st, _ = orc.render(u, imgs, tile_modulo=modulo, tile_remainder=i % modulo)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
            rays.append(st["rays"])
    total_t, total_r = sum(times), sum(rays)
    mrays = total_r / total_t / 1e6
    return {
        "value": round(mrays, 3), "unit": "Mrays/s", "cores": orc.threads, "kind": "port",
        "sample": f"1/{modulo} of the 16x16 tiles of each {w}x{h} frame (interleaved), {len(times)} frames, "
                  f"{total_r} rays in {total_t:.2f}s; SAH BVH build {build_s:.2f}s excluded",
        "_ms_per_step": 1e3 * total_t / len(times) * modulo, "_modulo": modulo,
    }


class Rig:
    """One process's share of the bench: context on the launching stream, torch.distributed plumbing, L2 flush buffer."""

    def __init__(self, world, rank, local_rank, exchange):
        import torch
        import torch.distributed as dist
        from metal4_raytracing_b200 import device
        self.torch, self.dist = torch, dist
        self.world, self.rank, self.local_rank, self.exchange = world, rank, local_rank, exchange
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        self.ctx = device.Context(local_rank)
        # one launching stream for the library, torch's L2 flush, the timing events and NCCL's stream ordering
        # (a non-default stream: handle 0 would mean "the context's own stream" to rt_set_stream)
        self.stream = torch.cuda.Stream(device=local_rank)
        torch.cuda.set_stream(self.stream)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t]

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(rig, workload, steps, warmup, e2e=True, sampler=None, verify=False, slice_of=1):
    """Device-timed frames (+ optionally the e2e loop and the N-rank == 1-rank frame check) of one workload.
    slice_of > 1 (tuning aid, single process): render only the tiles rank 0 of `slice_of` ranks would own."""
    torch = rig.torch
    from metal4_raytracing_b200 import _abi as A
    from metal4_raytracing_b200 import device, parallel, scene
    world, rank, ctx = rig.world, rig.rank, rig.ctx
    tile_world = slice_of if (world == 1 and slice_of > 1) else world
    sc, u, seeds, w, h = build_scene(workload)
    by_samples = world > 1 and rig.exchange == "samples"
    if by_samples:  # shares of the frame are summed in fp32; the motion-adaptive features need sample 0 on every rank
        u.enableMotionAdaptiveSampling = u.enableMotionAdaptiveAccumulation = 0
    rnd = device.Renderer(ctx, sc, w, h, seeds=seeds, fp32=by_samples)
    if workload == "K3env":
        rnd.set_environment(scene.procedural_sky(4096, 2048), 0.75, importance=True)
    xchg = parallel.FrameExchange(rnd, world, rank, mode=rig.exchange)
    animated = workload == "K5"
    pixels_owned = w * h if by_samples else int(parallel.owner_mask(w, h, tile_world, rank).sum())

    def partition():
        if by_samples:
            return xchg.draw_partition()
        return {"tile_modulo": tile_world, "tile_remainder": rank, "peers": xchg.peers_for_next_draw()}

    def frame(i, count=False):
        u.frameIndex = i
        if animated:
            sc.animate(i / 60.0)
            rnd.update()
        rnd.draw(u, count_rays=count, **partition())
        xchg.finish_frame()

    # ---- warm-up ------------------------------------------------------------------------------------------
    for i in range(warmup):
        frame(i)
    rig.barrier()
    if sampler is not None:
        sampler.start()
        time.sleep(0.3)
    frame(warmup)  # one more untimed frame: the GPU has idled while the clock sampler started
    # ---- timed region: device events on the launching stream, per-launch events for the roofline -----------
    launches0 = ctx.launches
    rnd.reset_ray_counters()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.kernel_timing(True)  # one event after each library launch, on the stream it was launched on
    rig.barrier()
    for k, i in enumerate(range(warmup + 1, warmup + 1 + steps)):
        if not os.environ.get("BENCH_NO_FLUSH"):  # diagnosis only: what the cold L2 costs a frame
            rig.flush.fill_(k & 0xFF)  # L2 flush, outside the per-frame events
        ev[k][0].record()
        frame(i, count="accumulate")  # ray counters: three warp-aggregated atomics per warp, always on
        ev[k][1].record()
    rig.barrier()
    frame_ms = [a.elapsed_time(b) for a, b in ev]
    ktimes = ctx.kernel_times()
    ctx.kernel_timing(False)
    counters = rnd.read_ray_counters()
    total_ms = float(sum(frame_ms))
    launches = ctx.launches - launches0
    if sampler is not None:
        sampler.stop()
    total_ms_max = rig.reduce([total_ms], "MAX")[0]
    rays_all, hits_all = rig.reduce([float(counters["rays"]), float(counters["hits"])], "SUM")
    res = {"workload": workload, "w": w, "h": h, "steps": steps, "frame_ms": frame_ms, "total_ms": total_ms,
           "total_ms_max": total_ms_max, "rays_all": rays_all, "hits_all": hits_all, "counters": counters,
           "ktimes": ktimes, "launches": launches, "pixels_owned": pixels_owned, "animated": animated,
           "mrays": rays_all / (total_ms_max * 1e-3) / 1e6, "e2e": None, "frame_equal": None}

    # ---- e2e through the host API: host inputs, D2H of the frame, wall clock -----------------------------------
    if e2e:
        rig.barrier()
        desc_bytes = 72 * sc.desc().instanceCount + 128 * sc.desc().lightCount + 208
        first = rnd.read_image(A.TEXTURE_ACCUMULATION)
        out_bytes = first.nbytes
        # the host-side frame buffers the results land in: two pinned frames, one read-back in flight (the reference
        # keeps up to three frames in flight, Renderer.swift:207) — the copy of frame i overlaps the rendering of
        # frame i + 1, which writes the other accumulation target
        host_frames = [ctx.pinned_array(first.shape, first.dtype) for _ in range(2)]
        rig.barrier()
        t0 = time.perf_counter()
        in_flight = None
        for k, i in enumerate(range(warmup + 1 + steps, warmup + 1 + 2 * steps)):
            u.frameIndex = i
            rig.flush.fill_(k & 0xFF)  # same L2 flush as the device-timed frames (here its 0.05 ms is inside the clock)
            if animated:
                sc.animate(i / 60.0)
            rnd.update()  # pinned H2D: instance descriptors, lights, (palettes); TLAS update
            rnd.draw(u, **partition())
            if in_flight is not None:
                ctx.download_wait(in_flight)  # frame i - 1 is on the host before anyone may overwrite its image
            xchg.finish_frame()
            if rank == 0:
                in_flight = rnd.read_image_async(A.TEXTURE_ACCUMULATION, host_frames[k & 1])  # D2H of the finished frame
        if in_flight is not None:
            ctx.download_wait(in_flight)
        rig.barrier()
        wall = rig.reduce([time.perf_counter() - t0], "MAX")[0]
        res["e2e"] = {"value": round(rays_all / wall / 1e6, 2), "unit": "Mrays/s",
                      "h2d_bytes_per_step": int(desc_bytes + (64 * 64 if animated else 0)),
                      "d2h_bytes_per_step": int(out_bytes), "ms_per_step": round(1e3 * wall / steps, 3),
                      "note": "ray count per frame taken from the device-timed frames (same workload, later sample "
                              "indices); the read-back of frame i (pinned host buffer, copy stream) overlaps the "
                              "rendering of frame i+1"}

    # ---- N ranks == 1 rank, bit for bit: two EMA frames from a cleared history, assembled on every rank
    # through the exchange; rank 0 then renders the same two frames alone and compares -------------------------------
    if verify and world > 1:
        def two_frames(modulo, remainder, exchange):
            rnd.reset_accumulation()
            if animated:  # same pose history on both renders: previous == current == pose 0 before frame 0
                sc.animate(0.0)
                rnd.update()
                rnd.update()
            for f in (0, 1):
                u.frameIndex = f
                if animated and f:
                    sc.animate(f / 60.0)
                    rnd.update()
                if exchange:
                    rnd.draw(u, **partition())
                    xchg.finish_frame()
                else:
                    rnd.draw(u, tile_modulo=modulo, tile_remainder=remainder)
        two_frames(world, rank, True)
        rig.barrier()
        assembled = rnd.read_image(A.TEXTURE_ACCUMULATION).copy() if rank == 0 else None
        rig.barrier()
        if rank == 0:
            two_frames(1, 0, False)
            ctx.sync()
            alone = rnd.read_image(A.TEXTURE_ACCUMULATION)
            if by_samples:  # float sums reassociated by the all-reduce: equal to rounding
                diff = float(np.abs(assembled.astype(np.float64) - alone).max() / max(1.0, float(np.abs(alone).max())))
                res["frame_equal"] = bool(diff <= 4e-6)
                res["frame_max_rel_diff"] = diff
            else:
                res["frame_equal"] = bool(np.array_equal(assembled.view(np.uint16), alone.view(np.uint16)))
            res["frame_sha256"] = {"assembled": hashlib.sha256(assembled.tobytes()).hexdigest()[:16],
                                   "one_rank": hashlib.sha256(alone.tobytes()).hexdigest()[:16]}
        rig.barrier()
    xchg.close()
    rnd.close()
    return res


def roofline_of(res):
    """The roofline object of the dominant kernel for one measured workload (this rank's launches)."""
    workload, steps, ktimes, c = res["workload"], res["steps"], res["ktimes"], res["counters"]
    rays, hits, closest_rays, shadow_rays = c["rays"], c["hits"], c["closest"], c["any"]
    total_ms = res["total_ms"]
    peak, peak_src = load_peaks()
    verts = 100000 if res["animated"] else 0
    b_ray = B_RAY[workload]
    # the persistent traversal kernel k_wf_traverse runs closest-hit rays and (fused into the next segment's launch) the
    # any-hit shadow rays; the library times its launches in the classes "trace" and "shadow"
    if "trace" in ktimes:
        trace_ms = ktimes["trace"][0] + ktimes.get("shadow", (0.0, 0))[0]
        trace_launches = ktimes["trace"][1] + ktimes.get("shadow", (0.0, 0))[1]
        dominant, dom_rays = "k_wf_traverse", closest_rays + shadow_rays
    else:
        trace_ms, trace_launches = ktimes.get("megakernel", (total_ms, steps))
        dominant, dom_rays = "k_trace_megakernel", rays
    achieved = dom_rays * b_ray / (trace_ms * 1e-3) / 1e9
    frame_bytes = rays * b_ray + hits * B_HIT + res["pixels_owned"] * B_PIXEL * steps + verts * B_VERTEX * steps
    kernels = {k: {"ms_per_step": round(v[0] / steps, 3), "launches_per_step": round(v[1] / steps, 1),
                   "share": round(v[0] / max(1e-9, total_ms), 4)} for k, v in ktimes.items()}
    if "shade" in ktimes:
        kernels["shade"]["achieved_gbs"] = round(hits * B_HIT / (ktimes["shade"][0] * 1e-3) / 1e9, 1)
    traffic, traffic_src = load_traffic(workload)
    roofline = {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_unit": "DRAM bytes per launch",
                "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": dominant, "launches": int(trace_launches),
                "avg_launch_ms": round(trace_ms / max(1, trace_launches), 4),
                "bytes_per_launch": round(dom_rays * b_ray / max(1, trace_launches)),
                "whole_frame": {"achieved": round(frame_bytes / (total_ms * 1e-3) / 1e9, 2),
                                "frac": round(frame_bytes / (total_ms * 1e-3) / 1e9 / peak, 4)},
                "kernels": kernels,
                "note": "algorithmic bytes = rays (closest-hit + any-hit) x %.0f B (SURVEY 8d) over the traversal kernel's "
                        "own launch time (classes trace + shadow: a launch traces segment k's closest hits and segment "
                        "k-1's shadow rays; with two pipeline lanes launches of different streams overlap at their "
                        "edges and each stretch of time is charged to the launch that ends it); whole_frame adds 300 B "
                        "per closest hit and 32 B per pixel over the frame time. The BVH fits L2 (traffic = measured "
                        "DRAM bytes per launch, profiles/), so the kernel is bound by thread-instruction issue "
                        "(`issue`), not by DRAM; `counted` uses the fetches a counter build counted instead of the "
                        "root-to-leaf model" % b_ray}
    # counted bytes per ray: node steps x 80 B + triangle tests x 48 B + instance entries x 64 B (tools/count_work.py)
    wc = _profile_json("work_counts.json").get(workload)
    if wc:
        per_ray = wc["bytes_per_ray"]
        roofline["counted"] = {"bytes_per_ray": round(per_ray, 1), "achieved": round(dom_rays * per_ray / (trace_ms * 1e-3) / 1e9, 2),
                               "frac": round(dom_rays * per_ray / (trace_ms * 1e-3) / 1e9 / peak, 4),
                               "nodes_per_ray": wc["nodes_per_ray"], "triangles_per_ray": wc["triangles_per_ray"],
                               "entries_per_ray": wc["entries_per_ray"], "source": "profiles/work_counts.json",
                               "note": "fetched bytes (mostly L1/L2 hits), so this may exceed what HBM could deliver"}
    iss = _profile_json("issue.json").get(workload)
    if iss:
        roofline["issue"] = iss
    return roofline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="K3", choices=sorted(WORKLOADS))
    ap.add_argument("--exchange", default="peer", choices=["peer", "gather", "samples"],
                    help="N > 1: frame assembly — tile partition with NVLink peer stores (default) or an NCCL all-gather, "
                         "or the sample partition with an NCCL all-reduce of rgba32f shares (parallel.py)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the short passes of the other configurations (N = 1)")
    ap.add_argument("--no-verify", action="store_true", help="skip the N-rank == 1-rank frame check (N > 1)")
    ap.add_argument("--slice", type=int, default=1, help="tuning aid (1 GPU): render only rank 0's tiles of this many "
                                                          "ranks; the line is marked and is not a bench result")
    args = ap.parse_args()
    steps, warmup = max(1, args.steps), max(3, args.warmup) if args.impl == "ours" else max(0, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": workload_text(args.workload),
              "l2": "new sample index every frame + 256 MiB L2 flush between timed frames",
              "sharding": ("single GPU" if world == 1 else "samples s % N == rank of every pixel, BVH replicated, NCCL all-reduce "
                           "of the rgba32f frame" if args.exchange == "samples" else "interleaved 16x16 tiles, BVH replicated"),
              "images": "rgba32f accumulation (shares are summed)" if (world > 1 and args.exchange == "samples")
                        else "rgba16f accumulation (reference format)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = run_oracle_sample(args.workload, steps, args.warmup, seconds_budget=60.0)
        line = {"metric": "Mrays/s", "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": args.warmup, "ms_per_step": round(cb.pop("_ms_per_step"), 3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "impl": "reference"}
        cb.pop("_modulo")
        line["cpu_baseline"] = cb
        line["e2e"] = {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        line["note"] = ("the reference is Swift/Metal and cannot run on Linux; this arm is the multithreaded C++ "
                        "port of its kernels (oracle/), ms_per_step extrapolated from the tile sample to a full frame")
        print(json.dumps(line), flush=True)
        return 0

    rig = Rig(world, rank, local_rank, args.exchange)
    sampler = ClockSampler(local_rank)
    res = measure(rig, args.workload, steps, warmup, e2e=not args.no_e2e, sampler=sampler, verify=not args.no_verify,
                  slice_of=args.slice)
    if args.slice > 1:
        config["slice"] = f"TUNING RUN: only the tiles of rank 0 of {args.slice} (no exchange); not a bench result"
    clocks = sampler.summary()
    roofline = roofline_of(res)

    others = None
    if world == 1 and not args.no_others and args.workload == "K3":
        # the other BASELINE configurations, three timed frames each, same process, outside the timed region above
        others = {}
        for name in OTHERS:
            try:
                r = measure(rig, name, 3, 3, e2e=True)
                rl = roofline_of(r)
                others[name] = {"workload": workload_text(name), "mrays": round(r["mrays"], 1),
                                "ms": round(r["total_ms_max"] / r["steps"], 3), "e2e": r["e2e"]["value"],
                                "e2e_ms": r["e2e"]["ms_per_step"], "rays_per_frame": int(r["rays_all"] / r["steps"]),
                                "roofline_frac": rl["frac"], "counted_frac": (rl.get("counted") or {}).get("frac"),
                                "issue_frac": (rl.get("issue") or {}).get("thread_instruction_frac"),
                                "kernels_ms": {k: v["ms_per_step"] for k, v in rl["kernels"].items()}}
            except Exception as exc:  # a missing asset directory must not cost the headline line
                others[name] = {"error": f"{type(exc).__name__}: {exc}"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_oracle_sample(args.workload, 3, 1, seconds_budget=20.0)
        cpu_baseline.pop("_ms_per_step")
        cpu_baseline.pop("_modulo")

    if rank == 0:
        line = {"metric": "Mrays/s", "value": round(res["mrays"], 2), "unit": "Mrays/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": round(res["total_ms_max"] / steps, 3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "rays_per_step": int(res["rays_all"] / steps), "frame_ms": [round(x, 3) for x in res["frame_ms"]],
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": res["e2e"],
                "gpu_launches": int(res["launches"]), "clocks": clocks,
                "exchange": args.exchange if world > 1 else None, "frame_equal": res["frame_equal"]}
        if res.get("frame_sha256"):
            line["frame_sha256"] = res["frame_sha256"]
        if "frame_max_rel_diff" in res:
            line["frame_max_rel_diff"] = res["frame_max_rel_diff"]
        if others is not None:
            line["others"] = others
        print(json.dumps(line), flush=True)
    rig.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
