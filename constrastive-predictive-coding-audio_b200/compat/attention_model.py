"""Reference module name ``attention_model`` (attention_model.py:9-82) -> cpc_b200."""
import _bootstrap  # noqa: F401
from cpc_b200.ar_models import AttentionModel, PositionalEncoder                                     # noqa: F401
