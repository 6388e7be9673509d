"""Mirror of the reference's internal/mip360.py entry the model uses (`new_space`, mip360.py:63-79)."""
from dataclasses import replace

from .mip import Samples


def new_space(samples: Samples) -> Samples:
    """Tags the handle; the contraction (mip360.py:47-60, threshold 0.1) and the linearised covariance update
    cov' = cov * v_j^2 (mip360.py:72-77) run inside the fused ray-march kernel."""
    return replace(samples, contracted=True)
