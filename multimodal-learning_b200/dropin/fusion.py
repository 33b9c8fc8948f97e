"""Drop-in for the reference's `fusion.py`: put this directory first on PYTHONPATH and the
reference's `from fusion import *` (networks_new.py:42) resolves to the B200 kernels."""
from multimodal_learning_b200.fusion import (BilinearFusion, PolynomialFusion, TrilinearFusion_A, TrilinearFusion_B,  # noqa: F401
                                              init_max_weights)
