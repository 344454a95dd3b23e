from prt_b200.mi_compat import *  # noqa: F401,F403
from prt_b200.mi_compat import ScalarPoint3f, ScalarVector3f, ScalarTransform4f  # noqa: F401
