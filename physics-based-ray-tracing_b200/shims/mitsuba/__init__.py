"""`import mitsuba as mi` -> the B200-backed subset in prt_b200.mi_compat (only used when the real
Mitsuba 3 wheel is not installed)."""
from prt_b200.mi_compat import *  # noqa: F401,F403
from prt_b200.mi_compat import (ad, warp, variant, variants, set_variant, register_integrator, register_sensor,  # noqa: F401
                                register_emitter, register_bsdf, load_dict, load_file, traverse, render)
from prt_b200 import mi_compat as _m

scalar_rgb = _m
llvm_ad_mono = _m
llvm_ad_rgb = _m
cuda_ad_mono = _m
cuda_ad_rgb = _m
