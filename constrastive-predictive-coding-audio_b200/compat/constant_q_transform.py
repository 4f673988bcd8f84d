"""Reference module name ``constant_q_transform`` (constant_q_transform.py:8-313) -> cpc_b200."""
import _bootstrap  # noqa: F401
from cpc_b200.frontend import (CQT, InverseCQT, PhaseAccumulation, PhaseDifference, abs, angle,      # noqa: F401
                               cqt_frequencies, pi, polar_to_complex, to_complex, unwrap)
