"""oracle/pyref.py -- TEST INFRASTRUCTURE: an independent pure-Python (binary64) transliteration of the reference's
scalar path loop ``_trace_single_ray`` (/root/reference/CustomIntegrator.py:262-376) and ``UltraBSDF.sample``
(/root/reference/CustomBSDF.py:87-175), written against tiny Mitsuba-like helpers for ``rectangle`` and ``sphere``
scenes only.  It exists to cross-check oracle/orc.c (the two share no code) and as the interpreter-speed proxy for the
reference's own execution model.  Canonical semantics: SURVEY.md Appendix F; the random
numbers are injected (same PCG32 / sample_tea_32 streams as everywhere else), because the reference is unseeded.
"""
from __future__ import annotations

import math

import numpy as np

RAY_EPS = 1500.0 * 2.0 ** -24
M64 = (1 << 64) - 1


class Pcg32:
    def __init__(self, seed: int, path: int):
        v0, v1 = self._tea((seed + (path >> 32)) & 0xffffffff, path & 0xffffffff)
        self.inc = ((v1 << 1) | 1) & M64
        self.state = 0
        self.next_u32()
        self.state = (self.state + v0) & M64
        self.next_u32()

    @staticmethod
    def _tea(v0, v1, rounds=4):
        s = 0
        for _ in range(rounds):
            s = (s + 0x9e3779b9) & 0xffffffff
            v0 = (v0 + ((((v1 << 4) & 0xffffffff) + 0xa341316c) ^ (v1 + s) ^ ((v1 >> 5) + 0xc8013ea4))) & 0xffffffff
            v1 = (v1 + ((((v0 << 4) & 0xffffffff) + 0xad90777d) ^ (v0 + s) ^ ((v0 >> 5) + 0x7e95761e))) & 0xffffffff
        return v0, v1

    def next_u32(self):
        old = self.state
        self.state = (old * 0x5851f42d4c957f2d + self.inc) & M64
        xs = (((old >> 18) ^ old) >> 27) & 0xffffffff
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xffffffff

    def next_f32(self) -> float:
        return float(np.uint32((self.next_u32() >> 9) | 0x3f800000).view(np.float32)) - 1.0


def _norm(v):
    return v / math.sqrt(float(v @ v))


def _frame(n):
    """Mitsuba coordinate_system."""
    sign = math.copysign(1.0, n[2])
    a = -1.0 / (sign + n[2])
    b = n[0] * n[1] * a
    return (np.array([(n[0] * n[0] * a) * sign + 1.0, b * sign, -n[0] * sign]),
            np.array([b, n[1] * n[1] * a + sign, -n[1]]))


class Shape:
    def __init__(self, kind, to_world, impedance, roughness):
        self.kind, self.M = kind, np.asarray(to_world, dtype=np.float64)
        self.Mi = np.linalg.inv(self.M)
        self.Z, self.rough = float(impedance), float(roughness)
        if kind == "sphere":
            self.c, self.r = self.M[:3, 3].copy(), float(np.linalg.norm(self.M[:3, 0]))
        else:
            self.n = _norm(self.Mi[:3, :3].T @ np.array([0.0, 0.0, 1.0]))

    def intersect(self, o, d, tmax):
        if self.kind == "sphere":
            oc = o - self.c
            A, B, C = d @ d, 2 * (oc @ d), oc @ oc - self.r ** 2
            disc = B * B - 4 * A * C
            if disc < 0:
                return None
            q = -0.5 * (B + math.copysign(math.sqrt(disc), B))
            t0, t1 = sorted((q / A, C / q))
            if not (t0 <= tmax and t1 >= 0) or (t0 < 0 and t1 > tmax):
                return None
            return t1 if t0 < 0 else t0
        ol = self.Mi[:3, :3] @ o + self.Mi[:3, 3]
        dl = self.Mi[:3, :3] @ d
        if dl[2] == 0:
            return None
        t = -ol[2] / dl[2]
        if not (0 <= t <= tmax) or abs(ol[0] + t * dl[0]) > 1 or abs(ol[1] + t * dl[1]) > 1:
            return None
        return t

    def interaction(self, o, d, t):
        if self.kind == "sphere":
            n = _norm(o + t * d - self.c)
            p = self.c + n * self.r
            loc = self.Mi[:3, :3] @ p + self.Mi[:3, 3]
            dp_du = self.M[:3, :3] @ np.array([-loc[1], loc[0], 0.0])
        else:
            ol = self.Mi[:3, :3] @ o + self.Mi[:3, 3]
            dl = self.Mi[:3, :3] @ d
            p = self.M[:3, :3] @ np.array([ol[0] + t * dl[0], ol[1] + t * dl[1], 0.0]) + self.M[:3, 3]
            n = self.n
            dp_du = self.M[:3, 0]
        s = dp_du - n * float(n @ dp_du)
        s = _norm(s) if float(s @ s) > 0 else _frame(n)[0]
        return p, n, s, np.cross(n, s)


def closest(shapes, o, d, tmax=math.inf):
    best, tb = None, tmax
    for i, sh in enumerate(shapes):
        t = sh.intersect(o, d, tb)
        if t is not None and (best is None or t < tb):
            best, tb = i, t
    return best, tb


def spawn(p, n, d):
    mag = (1.0 + float(np.max(np.abs(p)))) * RAY_EPS
    return p + n * math.copysign(mag, float(n @ d))


def ultra_bsdf(wi, n_g, n_s, Z, alpha, s1, s2):
    """CustomBSDF.py:87-175, literally (quirks Q4-Q9 of SURVEY.md Appendix A)."""
    fs, ft = _frame(n_g)                                            # :32
    w = np.array([wi @ fs, wi @ ft, wi @ n_g])                      # :33
    ws = _norm(np.array([alpha * w[0], alpha * w[1], w[2]]))        # :37-38
    inv = 1.0 / math.sqrt(max(1.0 - ws[2] * ws[2], 1e-7))           # :41
    T1 = np.array([ws[1] * inv, -ws[0] * inv, 0.0])                 # :42-44
    T2 = np.cross(ws, T1)                                           # :45
    r = 2.0 * s1 - 1.0                                              # :48 scalar sample -> disk diagonal
    qx = qy = 0.0 if r == 0 else r * math.cos(math.pi / 4)
    if r != 0:
        qy = r * math.sin(math.pi / 4)
    S = 0.5 * (1.0 + ws[2])                                         # :51
    qy = (1.0 - S) * math.sqrt(max(1.0 - qx * qx, 0.0)) + S * qy    # :52
    ms = qx * T1 + qy * T2 + math.sqrt(max(1.0 - qx * qx - qy * qy, 0.0)) * ws   # :55
    m = _norm(np.array([alpha * ms[0], alpha * ms[1], ms[2]]))      # :56-59
    if not (m @ wi < 0):                                            # :100
        m = -m
    cwm = float(wi @ m)                                             # :101
    Z1, Z2 = Z, 1.2                                                 # :104-107
    ratio = Z1 / Z2
    cTr = abs(cwm)
    sq = 1.0 - ratio ** 2 * (1.0 - cTr ** 2)                        # :120
    cTt = math.sqrt(max(sq, 0.0))
    Ar = (Z1 * cTr - Z2 * cTt) / (Z1 * cTr + Z2 * cTt)              # :123
    refl = wi + 2.0 * cwm * m                                       # :130
    trans = ratio * refl + (ratio * cTr - cTt) * m                  # :131
    reflect = sq < 0 or s2 < Ar * Ar                                # :137-145
    if reflect:
        return refl, 1.0 / (4.0 * abs(cwm)), Ar, True               # :154,170
    pdf_t = ratio ** 2 * abs(float(trans @ m)) / (abs(float(n_s @ wi)) * max(abs(float(n_s @ trans)), 1e-7))   # :158
    return trans, pdf_t, 1.0 - Ar, False


def acquire(shapes, params, seed=0, spp=1):
    """simulate_acquisition_parallel for every (angle, element, sample): returns (buf [n_a,n_e,T], tx [n_a,n_e],
    stats).  `params`: prt_b200.scene.AcqParams (duck-typed)."""
    n_a, n_e, T = params.n_angles, params.n_elements, params.time_samples
    c, fs, f = params.sound_speed, params.fs, params.frequency
    Tm = np.asarray(params.sensor_to_world, dtype=np.float64)
    nT = _norm(Tm[:3, :3] @ np.array([0.0, 0.0, 1.0]))
    a_m, a_c = math.radians(params.main_beam_deg), math.radians(params.cutoff_deg)
    buf = np.zeros((n_a, n_e, T))
    tx = np.zeros((n_a, n_e))
    st = dict(paths=0, segments=0, rays=0, deposits=0)
    ex = lambda e: params.pitch * (e - (n_e - 1) * 0.5)             # CI:84
    for a in range(n_a):
        th = math.radians(float(params.angles_deg[a]))
        for e in range(n_e):
            t0 = ex(e) * math.sin(th) / c                           # CI:87
            tx[a, e] = t0
            for s in range(spp):
                rng = Pcg32(seed, (a * n_e + e) * spp + s)
                o = Tm[:3, :3] @ np.array([ex(e), 0.0, 0.0]) + Tm[:3, 3]                      # CI:270-273
                d = _norm(Tm[:3, :3] @ np.array([math.sin(th), 0.0, math.cos(th)]))
                amp = atten = 1.0
                tof = geo = 0.0
                depth = 0
                st["paths"] += 1
                while depth < params.max_depth and geo < params.max_path_len:                 # CI:307
                    st["rays"] += 1
                    i, t = closest(shapes, o, d)                                              # CI:309
                    if i is None:
                        break
                    st["segments"] += 1
                    sh = shapes[i]
                    p, n, fs_, ft_ = sh.interaction(o, d, t)
                    geo += t
                    tof += t / c                                                              # CI:314-316
                    u_recv, s1, s2, u_rr = (rng.next_f32() for _ in range(4))                 # CI:319,337,365
                    recv = min(int(math.floor(u_recv * n_e)), n_e - 1)
                    tgt = Tm[:3, :3] @ np.array([ex(recv), 0.0, 0.0]) + Tm[:3, 3]
                    to_t = tgt - p
                    dist_recv = math.sqrt(float(to_t @ to_t))
                    sec = to_t / dist_recv                                                    # CI:322
                    st["rays"] += 1
                    visible = closest(shapes, spawn(p, n, sec), sec)[0] is None               # CI:324-325 (Q1)
                    atten *= math.exp(-params.attenuation * f * 1e-6 * t / 8.686)             # CI:328
                    Ttot = t0 + tof + dist_recv / c                                           # CI:329
                    phase = 2.0 * math.pi * f * Ttot                                          # CI:330
                    md = -d
                    wi = np.array([md @ fs_, md @ ft_, md @ n])
                    direction, pdf, a_resp, _ = ultra_bsdf(wi, n, n, sh.Z, sh.rough, s1, s2)  # CI:338
                    amp *= a_resp * float(n @ md) * max(pdf, 1e-6)                            # CI:340-341
                    al = abs(math.acos(max(-1.0, min(1.0, float(nT @ -sec)))))                # CI:292-295
                    w_i = 1.0 if al <= a_m else ((a_c - al) / (a_c - a_m) if al <= a_c else 0.0)
                    w_o = float(d @ n) / (n_a * n_e)                                          # CI:287,345
                    press = atten * amp * w_i * w_o * math.sin(phase)                         # CI:348
                    k = int(np.rint(Ttot * fs))                                               # CI:351-352
                    if 0 <= k < T and visible:
                        buf[a, recv, k] += press / spp                                        # CI:353-354
                        st["deposits"] += 1
                    d = _norm(direction)                                                      # CI:358-359 (Q9)
                    o = spawn(p, n, d)
                    depth += 1
                    rr = min(abs(atten * amp), 1.0)                                           # CI:364
                    survive = u_rr < rr
                    atten = atten / rr if survive else 0.0
                    if not survive or not (float(d @ nT) >= math.cos(a_c)):                   # CI:369-376
                        break
    return buf, tx, st


def shapes_from_desc(desc):
    out = []
    for s in desc.shapes:
        if s.kind not in ("sphere", "rectangle"):
            raise ValueError("pyref handles sphere / rectangle scenes only")
        m = desc.materials[s.material]
        out.append(Shape(s.kind, s.to_world, m.params[0], m.params[1]))
    return out


def pulse_shape(channel, fs, fc, sigma_s, cut=4.0):
    """Oracle for the "next" row f4 (pulse shaping): the prototype's echo model, /root/reference/RayTracingV0.py:185-204
    (`pulse(t, t0, amp, fc, sigma) = amp * sin(2 pi fc (t - t0)) * exp(-((t - t0)^2) / sigma^2)`, summed per echo),
    for echoes that sit on the sample grid: out[i] = sum_j in[j] h((i - j) / fs).  float64, truncated at |t| <= cut * sigma
    like the kernel (exp(-16) = 1.1e-7).  Test infrastructure only."""
    ch = np.asarray(channel, dtype=np.float64)
    half = int(math.ceil(cut * sigma_s * fs))
    n = np.arange(-half, half + 1, dtype=np.float64)
    h = np.sin(2.0 * np.pi * fc * n / fs) * np.exp(-((n / fs) ** 2) / sigma_s ** 2)
    flat = ch.reshape(-1, ch.shape[-1])
    out = np.empty_like(flat)
    for r in range(flat.shape[0]):
        out[r] = np.convolve(flat[r], h, mode="full")[half:half + flat.shape[1]]
    return out.reshape(ch.shape)


def das_beamform(ch, angles_deg, x, z, fs, c, pitch, t0=0.0, f_number=0.0):
    """Plane-wave delay-and-sum onto the pixel grid x (lateral) by z (depth): what the reference asks ultraspy for
    (`DelayAndSum.beamform`, /root/reference/USMain.py:175-200; ultraspy itself is not vendored -- restated from its
    published algorithm: linear interpolation of the channel data at t_tx(angle) + t_rx(element) - t0, optional
    f-number aperture, mean over the transmit angles).  float64, numpy; the checker of prt_das_beamform and the CPU
    figure beside bench.py's `us_render` entry."""
    n_a, n_e, T = ch.shape
    xe = pitch * (np.arange(n_e) - (n_e - 1) / 2)
    X, Z = np.meshgrid(x, z, indexing="ij")
    out = np.zeros_like(X, dtype=np.float64)
    for a in range(n_a):
        th = np.deg2rad(angles_deg[a])
        t_tx = (Z * np.cos(th) + X * np.sin(th)) / c
        for e in range(n_e):
            dx = X - xe[e]
            t = t_tx + np.sqrt(dx * dx + Z * Z) / c - t0
            s = t * fs
            i0 = np.floor(s).astype(np.int64)
            ok = (i0 >= 0) & (i0 + 1 < T)
            if f_number > 0:
                ok &= np.abs(dx) * 2 * f_number <= Z
            i0c = np.clip(i0, 0, T - 2)
            w = s - i0
            v = ch[a, e, i0c] * (1 - w) + ch[a, e, i0c + 1] * w
            out += np.where(ok, v, 0.0)
    return out / n_a
