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


def acquire_allreduce_pipelined(dev_scene, params, buf, tx, stats, stream, seed, spp_total, off, stride, dist, world, ps=None):
    """The acquisition of one rank + the sum all-reduce of its channel buffer, pipelined by steering angle: angle a's
    slice of ``buf`` is reduced (NCCL, on the process group's own stream) while angle a + 1 is still being traced, so
    the collective costs the step nothing but its last slice.  ``ps``: a prebuilt
    parameter struct (capi.make_acq_params), so the per-angle calls do no Python-side marshalling.  Returns after making
    ``stream`` wait for the reduces (nothing here blocks the host)."""
    if world <= 1:
        dev_scene.acquire_dev(params, buf.data_ptr(), tx.data_ptr() if tx is not None else 0, stats.data_ptr(), stream.cuda_stream,
                              seed=seed, spp=spp_total, sample_offset=off, sample_stride=stride, ps=ps)
        return
    works = []
    n_a = params.n_angles
    for a in range(n_a):
        dev_scene.acquire_dev(params, buf.data_ptr(), tx.data_ptr() if tx is not None else 0, stats.data_ptr(), stream.cuda_stream,
                              seed=seed, spp=spp_total, sample_offset=off, sample_stride=stride, angle_first=a, angle_count=1, ps=ps)
        works.append(dist.all_reduce(buf[a], op=dist.ReduceOp.SUM, async_op=True))
    for w in works:
        w.wait()


_STAT_CHUNKS, _STAT_BITS = 4, 20          # a counter (< 2^80) travels as 4 float32 words of 20 bits: sums over <= 16 ranks stay exact


def pack_stats(stats, shifts=None):
    """int64 tensor [8] -> float32 [8, 4]: 20-bit words, each exactly representable in binary32 and -- summed over up to 16
    ranks (< 2^24) -- still exact, so the counters can ride inside a float32 sum all-reduce."""
    import torch
    if shifts is None:
        shifts = torch.arange(_STAT_CHUNKS, dtype=torch.int64, device=stats.device) * _STAT_BITS
    return ((stats.view(-1, 1) >> shifts) & ((1 << _STAT_BITS) - 1)).to(torch.float32)


def unpack_stats(words) -> list:
    """Summed float32 words [8, 4] (numpy) -> the summed counters as Python ints (carries between words resolved here)."""
    w = np.asarray(words, dtype=np.float64).reshape(-1, _STAT_CHUNKS)
    return [int(sum(int(round(w[i, k])) << (_STAT_BITS * k) for k in range(_STAT_CHUNKS))) for i in range(w.shape[0])]


def acquire_sharded(dev_scene, params, seed: int, spp_total: int, to_host: bool = True, buf=None, stats=None):
    """Rank-local shard of an acquisition + sum all-reduce.  Returns (channel_buf, tx_delays, stats) with the buffers on
    the host (numpy, page-locked) if ``to_host`` else as torch CUDA tensors (owned by the scene's cache unless passed in:
    the next call overwrites them).

    Host path, per call: n_angles kernel launches and n_angles asynchronous all-reduces (angle a's slice is reduced while
    angle a + 1 is traced), ONE device-to-host copy and ONE stream synchronisation.  The path statistics are packed into 32
    floats right behind the channel buffer, so they are summed by the last slice's all-reduce and come back in the same
    copy (no second collective, no second copy); the transmit-delay table -- a function of the parameters alone -- is
    fetched once per parameter set and cached."""
    import torch
    from . import capi
    dist, rank, world = _dist()
    device = torch.device("cuda", dev_scene.ctx.device)
    off, stride, _ = shard_samples(spp_total, rank, world)
    shape = (params.n_angles, params.n_elements, params.time_samples)
    n = int(np.prod(shape))
    n_tail = 8 * _STAT_CHUNKS
    with torch.cuda.device(device):
        own = buf is None
        if own:
            flat = _cached(dev_scene, ("acq+tail", shape), lambda: torch.empty(n + n_tail, dtype=torch.float32, device=device))
            buf = flat[:n].view(shape)
        if stats is None:
            stats = _cached(dev_scene, "stats", lambda: torch.empty(8, dtype=torch.int64, device=device))
        stream = torch.cuda.current_stream(device)
        pkey = (params.n_angles, params.n_elements, params.time_samples, params.max_depth, params.pitch, params.fs, params.sound_speed,
                params.frequency, params.attenuation, params.main_beam_deg, params.cutoff_deg, params.max_path_len, int(params.quirk_flags),
                np.asarray(params.angles_deg, dtype=np.float64).tobytes(), np.asarray(params.sensor_to_world, dtype=np.float64).tobytes())
        cache = dev_scene.__dict__.setdefault("_shard_cache", {})
        ent = cache.get(("params", pkey))
        if ent is None:
            ent = cache[("params", pkey)] = {"ps": capi.make_acq_params(params), "tx_host": None,
                                             "tx": torch.empty(shape[:2], dtype=torch.float32, device=device)}
        tx = ent["tx"]
        want_tx = ent["tx_host"] is None
        if own:
            flat.zero_()
        else:
            buf.zero_()
        stats.zero_()
        if own and to_host:
            shifts = _cached(dev_scene, "stat_shifts", lambda: torch.arange(_STAT_CHUNKS, dtype=torch.int64, device=device) * _STAT_BITS)
            tail_view = flat[n:].view(8, _STAT_CHUNKS)
            last = flat[(params.n_angles - 1) * shape[1] * shape[2]:]          # last angle slice + the statistics words

            def pack():
                tail_view.copy_(pack_stats(stats, shifts))
            host = dev_scene.ctx.pinned_array((n + n_tail,), np.float32)
            host_t = torch.from_numpy(host)
            if world > 1:
                # three streams in flight: angle a + 1 is traced (this stream) while angle a's slice is summed over the ranks
                # (NCCL's stream) and angle a - 1's summed slice crosses PCIe into the page-locked result (copy stream)
                copy_stream = _cached(dev_scene, "copy_stream", lambda: torch.cuda.Stream(device=device))
                per = shape[1] * shape[2]
                for a in range(params.n_angles):
                    dev_scene.acquire_dev(params, buf.data_ptr(), tx.data_ptr() if want_tx else 0, stats.data_ptr(), stream.cuda_stream,
                                          seed=seed, spp=spp_total, sample_offset=off, sample_stride=stride, angle_first=a,
                                          angle_count=1, ps=ent["ps"])
                    final = a + 1 == params.n_angles
                    if final:
                        pack()
                    part = last if final else flat[a * per:(a + 1) * per]
                    w = dist.all_reduce(part, op=dist.ReduceOp.SUM, async_op=True)
                    with torch.cuda.stream(copy_stream):
                        w.wait()                                     # the COPY stream waits for the collective, not this one
                        host_t[a * per:a * per + part.numel()].copy_(part, non_blocking=True)
                if want_tx:
                    ent["tx_host"] = tx.cpu().numpy()                           # first call with these parameters only
                copy_stream.synchronize()
            else:
                dev_scene.acquire_dev(params, buf.data_ptr(), tx.data_ptr() if want_tx else 0, stats.data_ptr(), stream.cuda_stream,
                                      seed=seed, spp=spp_total, sample_offset=off, sample_stride=stride, ps=ent["ps"])
                pack()
                host_t.copy_(flat, non_blocking=True)
                if want_tx:
                    ent["tx_host"] = tx.cpu().numpy()
                stream.synchronize()
            hs = unpack_stats(host[n:])
            st = dict(paths=hs[0], segments=hs[1], rays=hs[2], deposits=hs[3], misses=hs[4])
            return host[:n].reshape(shape), ent["tx_host"], st
        acquire_allreduce_pipelined(dev_scene, params, buf, tx, stats, stream, seed, spp_total, off, stride, dist, world, ps=ent["ps"])
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
