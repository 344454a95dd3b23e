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


def acquire_sharded(dev_scene, params, seed: int, spp_total: int, to_host: bool = True, buf=None, stats=None):
    """Rank-local shard of an acquisition + all-reduce.  Returns (channel_buf, tx_delays, stats) with the
    buffers on the host (numpy) if ``to_host`` else as torch CUDA tensors."""
    import torch
    dist, rank, world = _dist()
    device = torch.device("cuda", dev_scene.ctx.device)
    off, stride, _ = shard_samples(spp_total, rank, world)
    shape = (params.n_angles, params.n_elements, params.time_samples)
    with torch.cuda.device(device):
        if buf is None:
            buf = torch.zeros(shape, dtype=torch.float32, device=device)
        else:
            buf.zero_()
        tx = torch.empty((params.n_angles, params.n_elements), dtype=torch.float32, device=device)
        if stats is None:
            stats = torch.zeros(8, dtype=torch.int64, device=device)
        else:
            stats.zero_()
        stream = torch.cuda.current_stream(device)
        dev_scene.acquire_dev(params, buf.data_ptr(), tx.data_ptr(), stats.data_ptr(), stream.cuda_stream, seed=seed,
                              spp=spp_total, sample_offset=off, sample_stride=stride)
        if world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)       # same stream: runs right behind the path kernel
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        if not to_host:
            return buf, tx, stats
        hb, htx, hs = buf.cpu().numpy(), tx.cpu().numpy(), stats.cpu().numpy()
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
            film = torch.zeros((rp.height, rp.width, 4), dtype=torch.float32, device=device)
        else:
            film.zero_()
        if stats is None:
            stats = torch.zeros(8, dtype=torch.int64, device=device)
        else:
            stats.zero_()
        stream = torch.cuda.current_stream(device)
        dev_scene.render_path_dev(rp, film.data_ptr(), stats.data_ptr(), stream.cuda_stream, seed=seed, spp=spp_total,
                                  sample_offset=off, sample_stride=stride)
        if world > 1:
            dist.all_reduce(film, op=dist.ReduceOp.SUM)
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        if develop:
            rgb = torch.empty((rp.height, rp.width, 3), dtype=torch.float32, device=device)
            dev_scene.develop_dev(film.data_ptr(), rp.height * rp.width, rgb.data_ptr(), stream.cuda_stream)
            film = rgb
        if not to_host:
            return film, stats
        hf, hs = film.cpu().numpy(), stats.cpu().numpy()
    return hf, dict(paths=int(hs[0]), segments=int(hs[1]), rays=int(hs[2]), shadow_rays=int(hs[3]))
