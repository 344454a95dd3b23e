"""Parity against fixtures produced by the REFERENCE'S OWN PYTHON (tests/golden/make_ref_fixtures.py).

tests/golden/ref_{bsdf,directivity,segments}.npz hold what /root/reference/CustomBSDF.py and
/root/reference/CustomIntegrator.py compute when they are executed unmodified (the Mitsuba / Dr.Jit calls they make
served by prt_b200.shims, ray queries by the oracle's intersector, uniforms injected from the per-path PCG32 streams).

  CPU  (-m "not gpu"): the oracle (binary32 and binary64) against those fixtures -- this is what pins the oracle --
                       and a regeneration check where /root/reference is mounted.
  GPU  (-m gpu):       the CUDA path through the C ABI (prt_ultra_bsdf_sample, prt_directivity_weights,
                       prt_acquire_trace, prt_acquire) against the same fixtures.

Decisions (primitive, receive element, visibility, reflect / transmit, time bin, number of segments) are compared
exactly; values with tolerances that follow the conditioning of the reference's own binary32 arithmetic: the phase is
~4e3 rad, so sin(phase) carries ~5e-4 absolute error in binary32, and `amp *= pdf` with
pdf = 1/(4 |wi.m|) (Q2, Q7) amplifies rounding without bound as wi.m -> 0.  Bars are therefore stated on quantiles.
"""
import json
import os
import sys

import numpy as np
import pytest

from conftest import HAS_REFERENCE, ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)


@pytest.fixture(scope="module")
def fx():
    return {k: np.load(os.path.join(GOLDEN, f"ref_{k}.npz")) for k in ("bsdf", "directivity", "segments")}


@pytest.fixture(scope="module")
def fixture_scenes():
    import make_ref_fixtures as M
    return dict(M.fixture_scenes())


def _rel(a, b, floor=1e-30):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    same_inf = (np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))) | (np.isnan(a) & np.isnan(b))   # inf * 0 on both sides
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(a - b) / np.maximum(np.abs(b), floor)
    r[same_inf] = 0.0
    return r


def _q(x, q):
    return float(np.quantile(x, q)) if len(x) else 0.0


# -------------------------------------------------------------------------------------------------
# fixture integrity
# -------------------------------------------------------------------------------------------------
def test_fixture_manifest(fx):
    meta = json.load(open(os.path.join(GOLDEN, "ref_fixtures.json")))
    assert meta["generator"] == "tests/golden/make_ref_fixtures.py"
    assert len(fx["bsdf"]["s1"]) == meta["n_bsdf"] >= 10000
    assert len(fx["directivity"]["w_o"]) == meta["n_directivity"]
    seg = fx["segments"]
    assert len(seg["scene_names"]) == 10 and int(seg["n_runs"]) == meta["n_runs"]
    for mode in "DP":
        r = seg["seg_" + mode]
        assert len(r) > 12000 and r["valid"].all()
        assert np.isin(np.arange(10), r["scene"]).all()
    # "D" is multi-bounce, "P" as written performs one segment per ray
    assert seg["seg_D"]["seg"].max() >= 1 and seg["seg_P"]["seg"].max() == 0
    if HAS_REFERENCE:
        import make_ref_fixtures as M
        for fn, h in meta["reference_files"].items():
            assert M.sha16(os.path.join(M.REFERENCE, fn)) == h, f"{fn} changed since the fixtures were generated"


@pytest.mark.skipif(not HAS_REFERENCE, reason="/root/reference not mounted (GPU box)")
def test_fixtures_regenerate_from_the_reference(fx):
    """The committed files ARE what the reference's code produces: re-run a slice of every part."""
    import make_ref_fixtures as M
    b = fx["bsdf"]
    sl = slice(0, 12000, 97)
    out = M.run_bsdf({k: b[k][sl] for k in ("wi", "ng", "ns", "impedance", "roughness", "s1", "s2")})
    for k, v in out.items():
        assert np.array_equal(v, b[k][sl], equal_nan=True), k
    d = M.run_directivity(60)
    full = fx["directivity"]
    # the generator draws its inputs from one stream, so a shorter run has different inputs: compare through the oracle's
    # contract instead -- weights in [0, 1], identical for the two reference implementations (D and P)
    assert np.array_equal(d["w_i_D"], d["w_i_P"]) and np.array_equal(full["w_i_D"], full["w_i_P"])
    seg = M.run_segments(quick=True, only=("Plate_Box:intended", "ring"))
    for mode in "DP":
        new = seg["seg_" + mode]
        old = fx["segments"]["seg_" + mode]
        old = old[(old["run"] == 0) & np.isin(old["scene"], np.unique(new["scene"]))]
        assert len(new) == len(old)
        for f in new.dtype.names:
            assert np.array_equal(new[f], old[f], equal_nan=True), (mode, f)


# -------------------------------------------------------------------------------------------------
# A. UltraBSDF.sample
# -------------------------------------------------------------------------------------------------
def _check_bsdf(b, d, pdf, amp, rf, what, bars):
    ref_rf = b["component"] == 0
    assert np.array_equal(ref_rf, b["sampled_type"] == 0x8)               # GlossyReflection <-> component 0 (CB:161-168)
    mism = np.flatnonzero(rf != ref_rf)
    # reflect iff TIR or s2 < Ar^2 (CB:137-145): a mismatch is only acceptable within rounding of that threshold
    ar = np.where(ref_rf, b["amp"], 1.0 - b["amp"]).astype(np.float64)
    if bars is not None:
        assert len(mism) <= bars["n_mismatch"], f"{what}: {len(mism)} reflect/transmit mismatches"
        assert np.all(np.abs(b["s2"][mism] - ar[mism] ** 2) < 1e-5), f"{what}: mismatch away from the threshold"
    ok = rf == ref_rf
    de = np.abs(d - b["dir"]).max(1) / np.maximum(np.linalg.norm(b["dir"], axis=1), 1.0)
    pe = _rel(pdf, b["pdf"])
    ae = np.abs(amp - b["amp"])
    stats = dict(dir=(float(np.median(de[ok])), _q(de[ok], 0.999), float(de[ok].max())),
                 pdf=(float(np.median(pe[ok])), _q(pe[ok], 0.999), float(pe[ok].max())),
                 amp=(float(np.median(ae[ok])), _q(ae[ok], 0.999), float(ae[ok].max())))
    for k, (med, p999, mx) in stats.items():
        assert bars is None or (med <= bars["med"] and p999 <= bars["p999"] and mx <= bars["max"]), \
            f"{what} {k}: med {med:.2e} p99.9 {p999:.2e} max {mx:.2e}"
    stats["n_mismatch"] = int(len(mism))
    return stats


def test_oracle_bsdf_matches_reference_python(fx, orc):
    b = fx["bsdf"]
    n = len(b["s1"])
    for prec, bars in ((32, dict(n_mismatch=0, med=1e-6, p999=2e-3, max=2e-2)), (64, dict(n_mismatch=2, med=1e-6, p999=5e-3, max=0.2))):
        d, pdf, amp, rf = orc.ultra_bsdf_n(b["wi"], b["ng"], b["ns"], b["impedance"], b["roughness"], b["s1"], b["s2"], prec)
        assert n >= 10000 and len(pdf) == n
        _check_bsdf(b, d, pdf, amp, rf, f"oracle f{prec}", bars)


@pytest.mark.gpu
def test_cuda_bsdf_matches_reference_python(fx):
    from prt_b200.engine import ultra_bsdf_sample
    b = fx["bsdf"]
    d, pdf, amp, rf = ultra_bsdf_sample(b["wi"], b["ng"], b["ns"], b["impedance"], b["roughness"], b["s1"], b["s2"])
    # measured on B200 (tools/ref_fixture_report.py, r02): 0 mismatches; median 1e-7; p99.9 dir 3.4e-4 / pdf 1.5e-3 / amp 2e-5;
    # max pdf 0.15 -- the tail is pdf = 1 / (4 |wi.m|) as wi.m -> 0, where one ulp of the dot product is a large relative step
    st = _check_bsdf(b, d, pdf, amp, rf, "cuda", dict(n_mismatch=3, med=2e-6, p999=5e-3, max=0.5))
    print("cuda bsdf vs reference python (median, p99.9, max):", st)


# -------------------------------------------------------------------------------------------------
# B. directivity weights
# -------------------------------------------------------------------------------------------------
def _groups(d):
    key = np.stack([d["sensor_to_world"][:, 3], d["main_beam_deg"], d["num_rays"]], -1)
    for k in np.unique(key, axis=0):
        yield np.flatnonzero(np.all(key == k, axis=1))


def test_oracle_directivity_matches_reference_python(fx, orc):
    d = fx["directivity"]
    assert np.array_equal(d["w_i_D"], d["w_i_P"])                         # the two implementations hold the same function
    assert 0.25 < np.mean(d["w_i_D"] > 0) < 0.75 and np.any((d["w_i_D"] > 0) & (d["w_i_D"] < 1))
    for prec, tol in ((32, 2e-5), (64, 2e-4)):      # binary64 differs from the reference's binary32 acos on the ramp
        wi, wo = orc.directivity_n(d["sensor_to_world"], d["sec_dir"], d["sec_dir"], d["normal"], d["main_beam_deg"],
                                   d["cutoff_deg"], d["num_rays"], prec)
        assert np.max(np.abs(wi - d["w_i_D"])) <= tol, (prec, np.max(np.abs(wi - d["w_i_D"])))
        assert np.max(np.abs(wo - d["w_o"]) * d["num_rays"]) <= 1e-6       # a dot product of unit vectors, / N (CI:117-118)


@pytest.mark.gpu
def test_cuda_directivity_matches_reference_python(fx):
    from prt_b200.engine import directivity_weights
    d = fx["directivity"]
    n_groups = 0
    for idx in _groups(d):
        wi, wo = directivity_weights(d["sensor_to_world"][idx[0]], d["sec_dir"][idx], d["sec_dir"][idx], d["normal"][idx],
                                     d["main_beam_deg"][idx[0]], d["cutoff_deg"][idx[0]], d["num_rays"][idx[0]])
        assert np.max(np.abs(wi - d["w_i_D"][idx])) <= 5e-5
        assert np.max(np.abs(wo - d["w_o"][idx]) * d["num_rays"][idx]) <= 1e-6
        n_groups += 1
    assert n_groups == 8


# -------------------------------------------------------------------------------------------------
# C / D. per-segment records of simulate_acquisition ("D") and simulate_acquisition_parallel ("P")
# -------------------------------------------------------------------------------------------------
def _flags(orc, mode):
    # "D" as literally written: tof of the last segment only, clamped bins, no abs in the roulette (Appendix A.1 #2,#3,#6);
    # "P" as literally written: one segment per ray (#8)
    return (orc.QF_TOF_LAST_SEGMENT | orc.QF_CLAMP_TIDX | orc.QF_RR_NO_ABS) if mode == "D" else orc.QF_SINGLE_BOUNCE


def _compare_segments(seg, names, scenes, mode, qf, trace, what, bars):
    """trace(desc, params, path_idx, seed, spp) -> records [n, max_depth] with the oracle's field names."""
    from prt_b200.scene import AcqParams
    r_all = seg["seg_" + mode]
    n_runs, seed = int(seg["n_runs"]), int(seg["seed"])
    report = {}
    for sid in np.unique(r_all["scene"]):
        name = str(names[sid])
        desc = scenes[name]
        p = AcqParams.from_props(desc.integrator, desc.sensor)
        p.quirk_flags = qf
        q = r_all[r_all["scene"] == sid]
        path = (q["a"].astype(np.uint64) * np.uint64(p.n_elements) + q["e"].astype(np.uint64)) * np.uint64(n_runs) + q["run"].astype(np.uint64)
        up, inv = np.unique(path, return_inverse=True)
        rec = trace(desc, p, up, seed, n_runs)
        o = rec[inv, q["seg"]]
        # --- the same number of segments per path ---------------------------------------------------
        n_ref = np.bincount(inv, minlength=len(up))
        n_ours = rec["valid"].sum(axis=1)
        bad_len = int(np.sum(n_ref != n_ours))
        # --- decisions ----------------------------------------------------------------------------
        dec = (o["valid"] == 1) & (o["prim"] == q["prim"]) & (o["shape"] == q["shape"]) & (o["recv"] == q["recv"]) & \
              (o["visible"] == q["visible"]) & (o["reflect"] == q["reflect"])
        dep = (q["deposited"] == 1) & dec
        kd = np.abs(o["k"].astype(np.int64) - q["k"])[dep]
        # --- values -------------------------------------------------------------------------------
        first = dec & (q["seg"] == 0)
        t_first = _rel(o["t"], q["t"])[first]
        t_all = _rel(o["t"], q["t"])[dec]
        # the reference itself overflows on a few records (usmain, 0 degrees: |n.wi| = 0 in binary32 -> pdf = inf, amp = inf,
        # press = inf * 0 = nan): values are compared where the reference is finite, and binary32 implementations must
        # be non-finite at exactly the same records
        fin_p, fin_a = np.isfinite(q["press"]), np.isfinite(q["amp"]) if mode == "D" else np.ones(len(q), bool)
        nonfinite = int(np.sum((np.isfinite(o["press"]) != fin_p)[dep]))
        pr = _rel(o["press"], q["press"], 1e-12)[dep & fin_p & (q["press"] != 0)]
        vals = dict(press=pr)
        if mode == "D":
            nonfinite += int(np.sum((np.isfinite(o["amp"]) != fin_a)[dec]))
            vals["amp"] = _rel(o["amp"], q["amp"], 1e-12)[dec & fin_a]
            vals["atten"] = _rel(o["atten"], q["atten"], 1e-12)[dec]
            vals["dir"] = np.abs(np.asarray(o["dir"], dtype=np.float64) - q["dir"]).max(1)[dec]
        report[name] = dict(n=len(q), nonfinite=nonfinite, bad_len=bad_len, bad_dec=int((~dec).sum()), k_off=int((kd > 0).sum()), k_max=int(kd.max()) if len(kd) else 0,
                            t_first_max=float(t_first.max()) if len(t_first) else 0.0, t_max=float(t_all.max()) if len(t_all) else 0.0,
                            **{k + "_med": float(np.median(v)) if len(v) else 0.0 for k, v in vals.items()},
                            **{k + "_p99": _q(v, 0.99) for k, v in vals.items()})
        rp = report[name]
        if bars is None:
            continue
        assert rp["bad_dec"] <= bars["bad_dec"] * len(q), f"{what} {mode} {name}: {rp}"
        assert rp["bad_len"] <= bars["bad_dec"] * len(up), f"{what} {mode} {name}: {rp}"
        assert rp["nonfinite"] <= bars.get("nonfinite", 0) * len(q), f"{what} {mode} {name}: {rp}"
        assert rp["k_max"] <= 1 and rp["k_off"] <= bars["k_off"] * max(len(kd), 1), f"{what} {mode} {name}: {rp}"
        assert rp["t_first_max"] <= bars["t_first"] and rp["t_max"] <= bars["t_any"], f"{what} {mode} {name}: {rp}"
        for k in vals:
            assert rp[k + "_med"] <= bars["med"] and rp[k + "_p99"] <= bars["p99"], f"{what} {mode} {name} {k}: {rp}"
    return report


ORACLE_BARS = {32: dict(bad_dec=0.0, k_off=0.0, t_first=1e-6, t_any=1e-4, med=1e-4, p99=2e-2),
               # binary64 against the reference's binary32: rounding-level decision flips become possible
               64: dict(nonfinite=1.0, bad_dec=0.002, k_off=0.002, t_first=2e-6, t_any=1e-4, med=3e-4, p99=5e-2)}


@pytest.mark.parametrize("mode", ["D", "P"])
def test_oracle_segments_match_reference_python(fx, fixture_scenes, orc, mode):
    seg = fx["segments"]
    for prec in (32, 64):
        def trace(desc, p, idx, seed, spp, prec=prec):
            return orc.OracleScene(desc).acquire_trace(p, idx, seed=seed, spp=spp, prec=prec)
        _compare_segments(seg, seg["scene_names"], fixture_scenes, mode, _flags(orc, mode), trace, f"oracle f{prec}", ORACLE_BARS[prec])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["D", "P"])
def test_cuda_segments_match_reference_python(fx, fixture_scenes, orc, mode):
    from prt_b200.engine import DeviceScene
    seg = fx["segments"]
    devs = {}

    def trace(desc, p, idx, seed, spp):
        dev = devs.setdefault(id(desc), DeviceScene(desc))
        return dev.acquire_trace(p, idx, seed=seed, spp=spp)
    # measured on B200 (r02, gpurun_out/ref_fixture_report.json -> profiles/r02_ref_fixture_report.json): NO decision differs on
    # any of the 32 000 records (bad_dec = bad_len = 0), one time bin off by one in 4 of 20 scene x mode cells, t <= 4.8e-7
    # (first segment), press median 2e-5 .. 1.8e-4 / p99 <= 2.1e-2, amp p99 <= 3e-2 (curved targets).  `nonfinite`: the
    # reference itself overflows on usmain's 0-degree transmission (pdf = inf); the CUDA path (fused multiply-adds) turns
    # non-finite on slightly different records there (34 of 1280)
    rep = _compare_segments(seg, seg["scene_names"], fixture_scenes, mode, _flags(orc, mode), trace, "cuda",
                            dict(nonfinite=0.03, bad_dec=0.002, k_off=0.004, t_first=2e-6, t_any=2e-4, med=5e-4, p99=5e-2))
    print(f"cuda vs reference python, {mode}:", json.dumps(rep))


# -------------------------------------------------------------------------------------------------
# the channel buffer / delay table the reference leaves behind (CustomIntegrator.py:43-46,203,257,260,354)
# -------------------------------------------------------------------------------------------------
def _compare_buffers(seg, names, scenes, mode, qf, acquire, what, tol_med, tol_p99):
    from prt_b200.scene import AcqParams
    n_runs, seed = int(seg["n_runs"]), int(seg["seed"])
    report = {}
    for sid, name in enumerate(names):
        desc = scenes[str(name)]
        p = AcqParams.from_props(desc.integrator, desc.sensor)
        p.quirk_flags = qf
        for run in range(n_runs):
            idx, val = seg[f"{mode}_{sid}_{run}_idx"], seg[f"{mode}_{sid}_{run}_val"].astype(np.float64)
            buf, tx = acquire(desc, p, seed, n_runs, run)
            got = np.asarray(buf, dtype=np.float64).reshape(-1) * n_runs     # ours averages over spp_total; one sample was traced
            if run == 0:
                assert np.allclose(np.asarray(tx).reshape(-1), seg[f"{mode}_{sid}_tx"], rtol=2e-6, atol=1e-13), (what, name)
            fin = np.isfinite(val)
            # support: every finite non-zero bin of the reference is non-zero here and vice versa (zero-valued deposits --
            # w_i = 0 -- leave no trace on either side)
            ours = np.flatnonzero((got != 0) | ~np.isfinite(got))
            extra, missing = np.setdiff1d(ours, idx), np.setdiff1d(idx, ours)
            rp = report.setdefault(str(name), dict(support_diff=0, med=0.0, p99=0.0))
            rp["support_diff"] += len(extra) + len(missing)
            assert tol_med is None or len(extra) + len(missing) <= 0.005 * max(len(idx), 1) + 1, (what, mode, name, run, len(extra), len(missing))
            if fin.sum() == 0:
                continue
            keep = fin & np.isfinite(got[idx])
            a, b = got[idx][keep], val[keep]
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-12)
            # quantiles, not an L2 norm: on curved targets single deposits carry amp ~ 1/|wi.m| (heavy tail, module docstring)
            rp["med"], rp["p99"] = max(rp["med"], float(np.median(rel))), max(rp["p99"], _q(rel, 0.99))
            assert tol_med is None or (np.median(rel) <= tol_med and _q(rel, 0.99) <= tol_p99), \
                (what, mode, name, run, float(np.median(rel)), _q(rel, 0.99))
    return report


@pytest.mark.parametrize("mode", ["D", "P"])
def test_oracle_buffers_match_reference_python(fx, fixture_scenes, orc, mode):
    seg = fx["segments"]

    def acquire(desc, p, seed, spp, run):
        buf, tx, _ = orc.OracleScene(desc).acquire(p, seed=seed, spp=spp, s_offset=run, s_stride=spp, prec=32, n_threads=4)
        return buf, tx
    _compare_buffers(seg, seg["scene_names"], fixture_scenes, mode, _flags(orc, mode), acquire, "oracle f32", 1e-4, 5e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["D", "P"])
def test_cuda_buffers_match_reference_python(fx, fixture_scenes, orc, mode):
    from prt_b200.engine import DeviceScene
    seg = fx["segments"]
    devs = {}

    def acquire(desc, p, seed, spp, run):
        dev = devs.setdefault(id(desc), DeviceScene(desc))
        buf, tx, _ = dev.acquire(p, seed=seed, spp=spp, sample_offset=run, sample_stride=spp)
        return np.array(buf), np.array(tx)
    _compare_buffers(seg, seg["scene_names"], fixture_scenes, mode, _flags(orc, mode), acquire, "cuda", 5e-4, 5e-2)   # measured: 2.1e-4, 3e-2
