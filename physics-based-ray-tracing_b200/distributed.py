"""Multi-GPU: sample-count shards + ONE sum all-reduce of the accumulator.

One process per GPU (torchrun); rank g traces samples s = g, g+G, g+2G, ... of every (angle, element) /
pixel with the SAME seed, so the union over ranks is exactly the 1-GPU path set; the scene and its BVH are
replicated (built redundantly per rank); the per-rank accumulator is a torch tensor so that
``torch.distributed.all_reduce`` (NCCL over NVLink/NVSwitch on the box; gloo in the CPU tests) can be
enqueued on the same stream right behind the path kernel.  torch is plumbing only: device memory, stream,
process group.  The reference has no counterpart (single process, CustomIntegrator.py:380-399).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_samples(spp_total: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(sample_offset, sample_stride, n_samples of this rank)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    n = (spp_total - rank + world - 1) // world if rank < spp_total else 0
    return rank, world, max(n, 0)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def allreduce_sum_(tensor):
    """In-place sum over ranks; no-op for world size 1."""
    dist, _, world = _dist()
    if world > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def _cached(dev_scene, key, make):
    """Per-scene cache of device tensors reused across calls (no allocator traffic in the steady state)."""
    cache = dev_scene.__dict__.setdefault("_shard_cache", {})
    t = cache.get(key)
    if t is None:
        t = cache[key] = make()
    return t


def _to_pinned(dev_scene, tensor):
    """Device tensor -> pooled page-locked numpy array (prt_host_alloc): one DMA, no pageable staging."""
    import torch
    host = dev_scene.ctx.pinned_array(tuple(tensor.shape), np.float32)
    torch.from_numpy(host).copy_(tensor)
    return host


def acquire_allreduce_pipelined(dev_scene, params, buf, tx, stats, stream, seed, spp_total, off, stride, dist, world):
    """The acquisition of one rank + the sum all-reduce of its channel buffer, pipelined by steering angle: angle a's
    slice of ``buf`` is reduced (NCCL, on the process group's own stream) while angle a + 1 is still being traced, so
    the collective costs the step nothing but its last slice.  Returns after making ``stream`` wait for the reduces."""
    if world <= 1:
        dev_scene.acquire_dev(params, buf.data_ptr(), tx.data_ptr(), stats.data_ptr(), stream.cuda_stream, seed=seed,
                              spp=spp_total, sample_offset=off, sample_stride=stride)
        return
    works = []
    for a in range(params.n_angles):
        dev_scene.acquire_dev(params, buf.data_ptr(), tx.data_ptr(), stats.data_ptr(), stream.cuda_stream, seed=seed,
                              spp=spp_total, sample_offset=off, sample_stride=stride, angle_first=a, angle_count=1)
        works.append(dist.all_reduce(buf[a], op=dist.ReduceOp.SUM, async_op=True))
    for w in works:
        w.wait()


def acquire_sharded(dev_scene, params, seed: int, spp_total: int, to_host: bool = True, buf=None, stats=None):
    """Rank-local shard of an acquisition + all-reduce.  Returns (channel_buf, tx_delays, stats) with the
    buffers on the host (numpy, page-locked) if ``to_host`` else as torch CUDA tensors (owned by the scene's cache
    unless passed in: the next call overwrites them)."""
    import torch
    dist, rank, world = _dist()
    device = torch.device("cuda", dev_scene.ctx.device)
    off, stride, _ = shard_samples(spp_total, rank, world)
    shape = (params.n_angles, params.n_elements, params.time_samples)
    with torch.cuda.device(device):
        if buf is None:
            buf = _cached(dev_scene, ("acq", shape), lambda: torch.empty(shape, dtype=torch.float32, device=device))
        buf.zero_()
        tx = _cached(dev_scene, ("tx", shape[:2]), lambda: torch.empty(shape[:2], dtype=torch.float32, device=device))
        if stats is None:
            stats = _cached(dev_scene, "stats", lambda: torch.empty(8, dtype=torch.int64, device=device))
        stats.zero_()
        stream = torch.cuda.current_stream(device)
        acquire_allreduce_pipelined(dev_scene, params, buf, tx, stats, stream, seed, spp_total, off, stride, dist, world)
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        if not to_host:
            return buf, tx, stats
        hb = _to_pinned(dev_scene, buf)
        htx, hs = tx.cpu().numpy(), stats.cpu().numpy()
    st = dict(paths=int(hs[0]), segments=int(hs[1]), rays=int(hs[2]), deposits=int(hs[3]), misses=int(hs[4]))
    return hb, htx, st


def render_sharded(dev_scene, rp, seed: int, spp_total: int, to_host: bool = True, film=None, stats=None, develop: bool = False):
    """Rank-local sample shard of a render + ONE all-reduce of the RGBW film.  ``develop``: divide by the weight
    channel on the device AFTER the reduce (so sharded == unsharded up to summation order) and return [H,W,3]."""
    import torch
    dist, rank, world = _dist()
    device = torch.device("cuda", dev_scene.ctx.device)
    off, stride, _ = shard_samples(spp_total, rank, world)
    with torch.cuda.device(device):
        if film is None:
            film = _cached(dev_scene, ("film", rp.height, rp.width),
                           lambda: torch.empty((rp.height, rp.width, 4), dtype=torch.float32, device=device))
        film.zero_()
        if stats is None:
            stats = _cached(dev_scene, "stats", lambda: torch.empty(8, dtype=torch.int64, device=device))
        stats.zero_()
        stream = torch.cuda.current_stream(device)
        dev_scene.render_path_dev(rp, film.data_ptr(), stats.data_ptr(), stream.cuda_stream, seed=seed, spp=spp_total,
                                  sample_offset=off, sample_stride=stride)
        if world > 1:
            dist.all_reduce(film, op=dist.ReduceOp.SUM)
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        if develop:
            rgb = _cached(dev_scene, ("rgb", rp.height, rp.width),
                          lambda: torch.empty((rp.height, rp.width, 3), dtype=torch.float32, device=device))
            dev_scene.develop_dev(film.data_ptr(), rp.height * rp.width, rgb.data_ptr(), stream.cuda_stream)
            film = rgb
        if not to_host:
            return film, stats
        hf, hs = _to_pinned(dev_scene, film), stats.cpu().numpy()
    return hf, dict(paths=int(hs[0]), segments=int(hs[1]), rays=int(hs[2]), shadow_rays=int(hs[3]))
