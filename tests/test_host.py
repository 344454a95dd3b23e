"""CPU: host-side logic -- transforms, XML / dict loaders, mesh readers, the C-ABI's exported symbols,
sample sharding and the gloo all-reduce path (world size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import HAS_REFERENCE, REFERENCE, ROOT
from prt_b200 import capi, scenes
from prt_b200.scene import AcqParams, load_dict_desc, load_xml
from prt_b200.transforms import Transform4f, apply_xml_ops

T = Transform4f


def test_transform_semantics():
    # look_at of every shipped sensor is the identity
    assert np.allclose(T().look_at([0, 0, 0], [0, 0, 0.05], [0, 1, 0]).matrix, np.eye(4))
    ops = [T().translate([0, 0, 0.08]), T().scale(0.06)]
    m = apply_xml_ops(ops, "mitsuba").matrix          # S @ T: translate first, then scale
    assert np.allclose(m[:3, 3], [0, 0, 0.0048]) and np.allclose(np.diag(m)[:3], 0.06)
    i = apply_xml_ops(ops, "intended").matrix         # T @ S
    assert np.allclose(i[:3, 3], [0, 0, 0.08])
    r = T().rotate([0, 1, 0], 90).matrix
    assert np.allclose(r[:3, :3] @ [0, 0, 1], [1, 0, 0], atol=1e-15)
    a = T().translate([1, 2, 3]) @ T().rotate([0, 0, 1], 30) @ T().scale([2, 3, 4])
    assert np.allclose((a @ a.inverse()).matrix, np.eye(4), atol=1e-12)
    assert np.allclose(a.transform_normal([0, 0, 1]) @ a.transform_vector([1, 0, 0]), 0, atol=1e-12)


def test_library_exports_every_declared_symbol():
    """Every function include/prt_b200.h declares is exported by libprt_b200.so (no compute calls here)."""
    with open(os.path.join(ROOT, "include", "prt_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    declared = set(re.findall(r"\b(prt_[a-z0-9_]+)\s*\(", text))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.prt_version().startswith(b"prt_b200")
    assert C.sizeof(capi.AcqParamsC) == 16 + 8 * 8 + 128 + 8 + 8
    assert capi.SEG_DTYPE.itemsize == 32 + 4 * 8


def test_ctypes_structs_match_the_header(tmp_path):
    """Every struct capi.py mirrors has the size and field offsets the C compiler gives include/prt_b200.h (ABI drift guard)."""
    pairs = {"prt_acq_params": capi.AcqParamsC, "prt_acq_stats": capi.AcqStatsC, "prt_bvh_stats": capi.BvhStatsC,
             "prt_render_params": capi.RenderParamsC, "prt_render_stats": capi.RenderStatsC, "prt_das_params": capi.DasParamsC,
             "prt_us_render_params": capi.UsRenderParamsC, "prt_kernel_times": capi.KernelTimesC}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "prt_b200.h")}"', "int main(void) {"]
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("prt_seg_record %zu\\n", sizeof(prt_seg_record));', "  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in pairs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)
    assert int(out["prt_seg_record"]) == capi.SEG_DTYPE.itemsize


def test_no_gpu_is_loud_not_a_fallback():
    """Without a CUDA device every compute entry point fails with a message; nothing computes on the CPU."""
    L = capi.load()
    n = C.c_int(-1)
    rc = L.prt_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert L.prt_create(0, C.byref(h)) != 0
    assert L.prt_last_error()
    from prt_b200.engine import DeviceScene
    with pytest.raises(capi.PrtError):
        DeviceScene(scenes.ultrasound_scene("Plate_Box"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "physics-based-ray-tracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, f), errors="replace") as fh:
                    src = fh.read()
                assert "orc_py" not in src and "liborc" not in src and "orc.h" not in src, os.path.join(dirpath, f)


def test_builtin_scenes():
    for name in scenes.MITSUBA_SCENES:
        for order in ("mitsuba", "intended"):
            d = scenes.ultrasound_scene(name, order)
            p = AcqParams.from_props(d.integrator, d.sensor)
            assert (p.n_angles, p.n_elements, p.time_samples, p.max_depth) == (5, 64, 10000, 10)
            assert p.fs == 50e6 and p.frequency == 3e6 and p.sound_speed == 1480 and p.pitch == 0.00012
            assert len(d.shapes) == (6 if scenes.MITSUBA_SCENES[name][1] else 1)
            assert all(d.materials[s.material].kind == "ultra" for s in d.shapes)
    v, vn, idx = scenes.ring_mesh()
    assert idx.shape == (1152, 3) and vn.shape == v.shape
    r = np.hypot(v[:, 0], v[:, 1])
    assert np.allclose(np.unique(np.round(r, 9)), [0.05, 0.06]) and np.allclose(np.unique(v[:, 2]), [0, 0.05])
    area = 0.5 * np.linalg.norm(np.cross(v[idx[:, 1]] - v[idx[:, 0]], v[idx[:, 2]] - v[idx[:, 0]]), axis=1).sum()
    assert abs(area - 0.04146) < 2e-4                 # total area of TestRing.obj
    c = scenes.cbox_scene(64, 4)
    assert c.n_triangles() == 12 and c.n_analytic() == 2
    assert [c.materials[s.material].kind for s in c.shapes[-2:]] == ["conductor", "dielectric"]
    assert np.allclose(c.materials[c.shapes[0].material].emission, 1.0)
    assert np.allclose(c.shapes[-2].to_world[:3, 3], [-0.3, -0.5, 0.2]) and np.isclose(c.shapes[-2].to_world[0, 0], 0.5)


@pytest.mark.skipif(not HAS_REFERENCE, reason="/root/reference is only present in the build container")
def test_reference_xml_equals_builtin_scenes():
    for name in scenes.MITSUBA_SCENES:
        for order in ("mitsuba", "intended"):
            x = load_xml(os.path.join(REFERENCE, "MitsubaScenes", name + ".xml"), transform_order=order)
            b = scenes.ultrasound_scene(name, order)
            assert len(x.shapes) == len(b.shapes)
            for sx, sb in zip(x.shapes, b.shapes):
                assert sx.kind == sb.kind and np.allclose(sx.to_world, sb.to_world, atol=1e-15)
                assert np.allclose(x.materials[sx.material].params, b.materials[sb.material].params)
            px, pb = AcqParams.from_props(x.integrator, x.sensor), AcqParams.from_props(b.integrator, b.sensor)
            for f in ("n_elements", "pitch", "time_samples", "max_depth", "fs", "sound_speed", "frequency", "attenuation",
                      "main_beam_deg", "cutoff_deg"):
                assert getattr(px, f) == getattr(pb, f), f
            assert np.array_equal(px.angles_deg, pb.angles_deg) and np.allclose(px.sensor_to_world, pb.sensor_to_world)


@pytest.mark.skipif(not HAS_REFERENCE, reason="/root/reference is only present in the build container")
def test_reference_cbox_and_meshes_load():
    x = load_xml(os.path.join(REFERENCE, "scenes", "cbox.xml"), res=64, spp=4)
    b = scenes.cbox_scene(64, 4)
    assert x.n_triangles() == b.n_triangles() == 12 and x.n_analytic() == b.n_analytic() == 2
    assert int(x.film["width"]) == 64 and int(x.sampler["sample_count"]) == 4 and x.rfilter.plugin_name() == "tent"
    for sx, sb in zip(x.shapes, b.shapes):
        assert sx.kind == sb.kind and np.allclose(sx.to_world, sb.to_world)
        if sx.kind == "mesh":
            wx = sx.v[sx.idx.reshape(-1)]
            wb = sb.v[sb.idx.reshape(-1)]
            assert np.allclose(wx, wb)
        assert x.materials[sx.material].kind == b.materials[sb.material].kind
        assert np.allclose(x.materials[sx.material].params, b.materials[sb.material].params)
        assert np.allclose(x.materials[sx.material].emission, b.materials[sb.material].emission)
    from prt_b200.meshio import load_mesh
    v, vn, idx = load_mesh(os.path.join(REFERENCE, "TestRing", "TestRing.obj"))
    assert idx.shape == (1152, 3) and vn is not None and v.shape == vn.shape
    for name, nt, has_n in (("teapot.ply", 2256, False), ("bunny.ply", 69451, False), ("suzanne.ply", 62976, True),
                            ("ico_10k.ply", 20480, True)):
        v, vn, idx = load_mesh(os.path.join(REFERENCE, "scenes", "meshes", name))
        assert idx.shape == (nt, 3) and (vn is not None) == has_n and idx.max() < len(v)


@pytest.mark.skipif(not HAS_REFERENCE, reason="/root/reference is only present in the build container")
def test_reference_scripts_run_unchanged_up_to_the_gpu_boundary():
    """The reference's own scripts, byte for byte, through tools/run_reference_script.py: TestScene.py completes;
    USMain.py gets through its imports, plugin registration, scene dict and mi.load_dict (USMain.py:1-259) and -- in this
    GPU-less container -- stops at the first compute call (simulate_acquisition_parallel, :99) with a loud error, not
    with a CPU fallback.  (With a GPU it runs to the end; that call sequence is tests/test_gpu_driver.py.)"""
    run = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py")]
    r = subprocess.run(run + [os.path.join(REFERENCE, "TestScene.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    n = C.c_int(-1)
    if capi.load().prt_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present: USMain.py would run its 51 acquisitions")
    r = subprocess.run(run + [os.path.join(REFERENCE, "USMain.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "USMain.py\", line 99, in us_render" in r.stderr and "simulate_acquisition_parallel" in r.stderr
    assert "PrtError" in r.stderr and ("no CPU fallback" in r.stderr or "CUDA" in r.stderr)
    assert "ModuleNotFoundError" not in r.stderr and "KeyError" not in r.stderr and "AttributeError" not in r.stderr


def test_scene_parameters_expose_shape_transforms():
    """mi.traverse(scene) lists '<shape id>.to_world' next to the BSDF keys; before the scene is on a device an edit lands in
    the scene description (on a device it becomes a refit: tests/test_gpu_edges.py)."""
    from prt_b200 import mi_compat as mi
    from prt_b200 import scenes
    from prt_b200.transforms import Transform4f
    scene = mi.Scene(scenes.test_ring_scene())
    params = mi.traverse(scene)
    assert "ring.to_world" in params and "ring.bsdf.roughness" in params
    T = Transform4f().translate([0.0, 0.0, 0.01]) @ Transform4f(scene.desc.shapes[0].to_world)
    params["ring.to_world"] = T
    assert params.update() == ["ring.to_world"]
    assert np.allclose(scene.desc.shapes[0].to_world, T.matrix)
    with pytest.raises(KeyError):
        params["nosuch.to_world"] = T


def test_packed_statistics_survive_a_float32_sum():
    """distributed.acquire_sharded sums its u64 path counters inside the float32 all-reduce of the channel buffer: packed as
    20-bit words they must come back exact for up to 16 ranks."""
    import torch
    from prt_b200.distributed import pack_stats, unpack_stats
    g = torch.Generator().manual_seed(1)
    ranks = [torch.randint(0, 2 ** 62, (8,), generator=g, dtype=torch.int64) for _ in range(16)]
    ranks[3][0] = 2 ** 62 - 1
    ranks[5][1] = 0
    acc = torch.zeros(8, 4, dtype=torch.float32)
    for r in ranks:
        acc += pack_stats(r)                          # what the float32 sum all-reduce computes
    want = [sum(int(r[i]) for r in ranks) for i in range(8)]
    assert unpack_stats(acc.numpy()) == want


def test_mesh_readers_roundtrip(tmp_path):
    from prt_b200.meshio import load_obj, load_ply
    p = tmp_path / "q.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\nf -4 -3 -2\n")
    v, vn, idx = load_obj(str(p))
    assert idx.tolist()[:2] == [[0, 1, 2], [0, 2, 3]] and vn is not None and len(idx) == 3
    ascii_ply = tmp_path / "a.ply"
    ascii_ply.write_text("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
                         "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n")
    v, vn, idx = load_ply(str(ascii_ply))
    assert v.shape == (4, 3) and idx.tolist() == [[0, 1, 2], [0, 2, 3]]
    import struct
    binary = tmp_path / "b.ply"
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
           "property float nx\nproperty float ny\nproperty float nz\nelement face 1\nproperty list uchar uint vertex_indices\nend_header\n")
    body = b"".join(struct.pack("<6f", *r) for r in ((0, 0, 0, 0, 0, 1), (1, 0, 0, 0, 0, 1), (0, 1, 0, 0, 0, 1)))
    body += struct.pack("<B3I", 3, 0, 1, 2)
    binary.write_bytes(hdr.encode() + body)
    v, vn, idx = load_ply(str(binary))
    assert v.shape == (3, 3) and np.allclose(vn, [[0, 0, 1]] * 3) and idx.tolist() == [[0, 1, 2]]


def test_plugin_surface_and_scene_parameters():
    """The Python surface USMain.py touches (no GPU call is made: nothing is traced here)."""
    from prt_b200 import shims
    shims.install()
    import mitsuba as mi
    import drjit as dr
    mi.set_variant("llvm_ad_mono")
    mi.set_variant("cuda_ad_mono")
    with pytest.raises(AttributeError):
        mi.set_variant("no_such_variant")
    from CustomIntegrator import UltraIntegrator
    from CustomSensor import UltraSensor, CustomSensor  # noqa: F401
    from CustomEmmitter import CustomEmitter
    from CustomBSDF import UltraBSDF
    mi.register_integrator("ultrasound_integrator", UltraIntegrator)
    mi.register_sensor("ultrasound_sensor", UltraSensor)
    mi.register_emitter("ultrasound_emitter", CustomEmitter)
    mi.register_bsdf("ultrasound_bsdf", UltraBSDF)
    d = scenes.usmain_scene_dict()
    d["integrator"]["angles"] = dr.linspace(mi.Float, -15, 15, 5)
    scene = mi.load_dict(d)
    integ = scene.integrator()
    assert isinstance(integ, UltraIntegrator) and isinstance(scene.sensors()[0], UltraSensor)
    assert (integ.n_angles, integ.n_elements, integ.time_samples, integ.max_depth) == (5, 64, 10000, 10)
    assert np.allclose(integ.angles.numpy(), [-15, -7.5, 0, 7.5, 15]) and integ.fs == 50e6 and integ.frequency == 5e6
    assert np.allclose(scene.sensors()[0].transform.matrix, np.eye(4))
    col, active, aov = integ.sample(scene, None, None, None, True)
    assert float(col[0]) == 0.0 and aov == []
    params = mi.traverse(scene)
    assert "flat_plate.bsdf.roughness" in params and "wall_back.bsdf.impedance" in params
    params["shape.bsdf.roughness"] = 0.25           # the key the driver writes (USMain.py:264)
    assert params.update() == ["shape.bsdf.roughness"]
    assert all(m.params[1] == 0.25 for m in scene.desc.materials)
    with pytest.raises(KeyError):
        params["nonexistent.key"] = 1.0
    # default-constructed plugins carry the reference's defaults (CustomIntegrator.py:16-42, CustomBSDF.py:12-18)
    from prt_b200.scene import Properties
    u = UltraIntegrator(Properties("ultrasound_integrator"))
    assert (u.max_depth, u.n_elements, u.n_angles, u.time_samples, u.pitch) == (2, 128, 25, 3000, 0.00035)
    b = UltraBSDF(Properties("ultrasound_bsdf"))
    assert abs(float(b.impedance[0]) - 1.54) < 1e-6 and float(b.roughness[0]) == 0.5
    assert b.eval(None, None, None, True) == 0.0 and b.eval_pdf(None, None, None, True) == (0.0, 0.0)
    e = CustomEmitter(Properties("ultrasound_emitter"))
    ps, pdf = e.sample_position(0.0, (0.5, [0.5, 0.5]))
    assert abs(pdf - 1 / (64 * 0.0003 * 0.0005)) < 1e-6 * pdf
    ray, w = e.sample_ray(0.0, 0.999, [0.5, 0.5], 0.5)
    assert np.allclose(np.asarray(ray.d), [0, 0, 1], atol=1e-7)


def test_shard_samples_partition():
    from prt_b200.distributed import shard_samples
    for spp in (1, 7, 64, 209716):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                off, stride, n = shard_samples(spp, r, world)
                s = list(range(off, spp, stride))
                assert len(s) == n
                seen += s
            assert sorted(seen) == list(range(spp))


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from prt_b200.distributed import allreduce_sum_, shard_samples
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
off, stride, n = shard_samples(9, rank, 2)
# each rank deposits 1.0 into bin s for every sample s it owns: the all-reduced buffer must be all ones
buf = torch.zeros(9)
for s in range(off, 9, stride):
    buf[s] += 1.0
allreduce_sum_(buf)
assert torch.equal(buf, torch.ones(9)), buf
st = torch.tensor([n], dtype=torch.int64)
allreduce_sum_(st)
assert int(st[0]) == 9
# the per-angle pipelined reduce (acquire_allreduce_pipelined): a stand-in scene fills angle a's slice when its launch
# is requested; every slice must come back summed over ranks, whatever order the asynchronous reduces complete in
from prt_b200.distributed import acquire_allreduce_pipelined
class _Params: n_angles = 5
class _Stream: cuda_stream = 0
class _FakeScene:
    def __init__(self, buf): self.buf, self.calls = buf, []
    def acquire_dev(self, params, buf_ptr, tx_ptr, stats_ptr, stream, seed=0, spp=1, sample_offset=0, sample_stride=1,
                    angle_first=0, angle_count=None, ps=None):
        self.calls.append((angle_first, angle_count, sample_offset, sample_stride))
        self.buf[angle_first] += float(10 * (rank + 1) + angle_first)
pbuf = torch.zeros(5, 7)
fake = _FakeScene(pbuf)
acquire_allreduce_pipelined(fake, _Params(), pbuf, torch.zeros(1), torch.zeros(1), _Stream(), 0, 4, rank, 2, dist, 2)
assert fake.calls == [(a, 1, rank, 2) for a in range(5)], fake.calls
for a in range(5):
    assert torch.equal(pbuf[a], torch.full((7,), float(10 + 20 + 2 * a))), (a, pbuf[a])
# acquire_sharded's layout: the path counters ride as 32 float words right behind the LAST angle slice, in ITS all-reduce
from prt_b200.distributed import pack_stats, unpack_stats
flat = torch.zeros(5 * 7 + 32)
flat[4 * 7:5 * 7] = float(rank + 1)
mine = torch.tensor([2 ** 40 + rank, 3, 5 * (rank + 1), 0, 2 ** 33 * (rank + 1), 0, 0, 1], dtype=torch.int64)
flat[5 * 7:].view(8, 4).copy_(pack_stats(mine))
dist.all_reduce(flat[4 * 7:], op=dist.ReduceOp.SUM)
assert torch.equal(flat[4 * 7:5 * 7], torch.full((7,), 3.0))
assert unpack_stats(flat[5 * 7:].numpy()) == [2 ** 41 + 1, 6, 15, 0, 3 * 2 ** 33, 0, 0, 2]
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world_size_2_allreduce(tmp_path):
    """The N>1 host path (shard -> accumulate -> ONE sum all-reduce) on CPU with the gloo backend."""
    w = tmp_path / "worker.py"
    w.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(w), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bench_line_contract_helpers():
    """bench.py pieces that do not need a GPU: both arms build the SAME `config` object; the roofline of an on-chip scene is an
    instruction-issue fraction that can never exceed 1 (a stale per-ray instruction count yields `stale`, not a number); the
    HBM-resident scene reports an HBM fraction with the measured DRAM traffic beside it; an `also` entry stays small enough for
    the four of them to survive a 1 500-character tail."""
    import json
    import bench
    from prt_b200.scene import AcqParams
    desc, label = bench.workload_desc(bench.DEFAULT_WORKLOAD)
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    a = bench.acq_config(label, p, bench.C2_SPP, 1, desc.n_triangles(), desc.n_analytic())
    b = bench.acq_config(label, p, bench.C2_SPP, 1, desc.n_triangles(), desc.n_analytic())
    assert a == b and a["paths_per_gpu_per_step"] == 67109120 and "intended" in a["workload"]
    clk = {"sm_mhz": 1965.0}
    e = bench.ncu_entry(bench.DEFAULT_WORKLOAD)
    assert e and e["kernel"] == "prt::k_acquire<false>" and 10 < e["warp_inst_per_ray"] < 40
    rays = 53.7e6
    rf = bench.roofline_block(bench.DEFAULT_WORKLOAD, "prt::k_acquire<false>", 1.1, rays, 2.56e6, clk, on_chip=True)
    assert rf["bound"] == "issue" and 0.3 < rf["frac"] <= 1.0 and rf["unit"] == "Gwarp-inst/s" and rf["hbm"]["dram_frac"] < 1e-3
    fast = bench.roofline_block(bench.DEFAULT_WORKLOAD, "prt::k_acquire<false>", 0.1, rays, 2.56e6, clk, on_chip=True)   # impossible rate
    assert fast["frac"] is None and fast["stale"] is True
    other = bench.roofline_block(bench.DEFAULT_WORKLOAD, "prt::some_other_kernel", 1.1, rays, 2.56e6, clk, on_chip=True)
    assert other["frac"] is None and other["traffic"] is None            # a capture of a different kernel is not evidence
    hf = bench.roofline_block("heightfield", "prt::k_wf_trace<false>", 9.3, 16.6e6, 16.6e6 * 1728, clk, on_chip=False)
    assert hf["bound"] == "hbm" and 0.3 < hf["frac"] < 1.0 and 0.05 < hf["hbm"]["dram_frac"] < hf["frac"]
    line = {"value": 48622.4, "msamples_per_s": 12400.4, "ms_per_step": 5.41, "n_gpus": 8, "kernel": "prt::k_wf_trace<false>",
            "e2e": {"value": 47188.0, "host_checksum": 0.00306}, "roofline": hf, "cpu_baseline": {"value": 104.0, "cores": 16}}
    c = bench.compact(line)
    assert len(json.dumps(c)) < 300 and c["bound"] == "hbm" and c["n"] == 8
    assert bench.is_pt("cbox") and bench.is_pt("heightfield:708") and not bench.is_pt("ring")
