"""``UltraBSDF`` -- drop-in for /root/reference/CustomBSDF.py:7-191, evaluated on the GPU.

Same constructor properties (impedance -> 1.54, roughness -> 0.5; :12-18), same flags (:22-26), same
``sample(ctx, si, sample1, sample2, active) -> (BSDFSample3f, amplitude)`` (:87-175), ``eval/pdf -> 0``,
``eval_pdf -> (0, 0)`` (:177-184) and ``traverse`` keys (:186-188).  ``sample`` runs the batched
``prt_ultra_bsdf_sample`` kernel; inside an acquisition the same device function is inlined in the
path kernel, so this entry point exists for API parity and tests.
"""
import numpy as np

from prt_b200 import mi_compat as mi


class UltraBSDF(mi.BSDF):
    _prt_material_kind = "ultra"

    def __init__(self, props):
        super().__init__(props)
        self.impedance = mi.Float(1.54)
        if props.has_property('impedance'):
            self.impedance = mi.Float(props['impedance'])
        self.roughness = mi.Float(0.5)
        if props.has_property('roughness'):
            self.roughness = mi.Float(props['roughness'])
        reflection_flags = mi.BSDFFlags.DeltaReflection | mi.BSDFFlags.FrontSide | mi.BSDFFlags.BackSide
        transmission_flags = mi.BSDFFlags.DeltaTransmission | mi.BSDFFlags.FrontSide | mi.BSDFFlags.BackSide
        self.m_components = [reflection_flags, transmission_flags]
        self.m_flags = reflection_flags | transmission_flags

    def flags(self):
        return self.m_flags

    def sample(self, ctx, si, sample1, sample2, active=True):
        from prt_b200.engine import ultra_bsdf_sample
        wi = np.asarray(si.wi, dtype=np.float32).reshape(-1, 3)
        ng = np.asarray(si.n, dtype=np.float32).reshape(-1, 3)
        ns = np.asarray(si.sh_frame.n, dtype=np.float32).reshape(-1, 3)
        n = wi.shape[0]
        s1 = np.broadcast_to(np.asarray(sample1, dtype=np.float32).reshape(-1), (n,))
        # CustomBSDF.py:144 collapses the reflect/transmit test to lane 0 (Q10); with width-1 calls, as the
        # reference makes them, that is the per-lane test
        s2 = np.broadcast_to(np.asarray(sample2, dtype=np.float32).reshape(-1), (n,))
        d, pdf, amp, refl = ultra_bsdf_sample(wi, ng, ns, float(np.asarray(self.impedance).reshape(-1)[0]),
                                              float(np.asarray(self.roughness).reshape(-1)[0]), s1, s2)
        bs = mi.BSDFSample3f()
        bs.sampled_type = mi.UInt32(np.where(refl, int(mi.BSDFFlags.GlossyReflection), int(mi.BSDFFlags.GlossyTransmission)))
        # CB:165  bs.wo = si.to_local(chosen_dir); the integrator's si.to_world(bs.wo) undoes it (Q9)
        bs.wo = si.to_local(d) if hasattr(si, "to_local") else mi.Vector3f(d)
        bs.pdf = mi.Float(pdf)
        bs.eta = mi.Float(1.0)
        bs.sampled_component = mi.UInt32(np.where(refl, 0, 1))
        return (bs, mi.Float(amp))

    def eval(self, ctx, si, wo, active=True):
        return 0.0

    def pdf(self, ctx, si, wo, active=True):
        return 0.0

    def eval_pdf(self, ctx, si, wo, active=True):
        return 0.0, 0.0

    def traverse(self, callback):
        callback.put_parameter('impedance', self.impedance, mi.ParamFlags.Differentiable)
        callback.put_parameter('roughness', self.roughness, mi.ParamFlags.Differentiable)

    def parameters_changed(self, keys=None):
        pass
