"""Reference module name ``audio_model`` (audio_model.py:8-307) -> cpc_b200."""
import _bootstrap  # noqa: F401
from cpc_b200.encoders import AudioEncoder, encoder_default_dict                                     # noqa: F401
from cpc_b200.ar_models import AudioGRUModel, ConvolutionalArBlock, ConvolutionalArModel             # noqa: F401
from cpc_b200.model import (ActivationRegister, ActivationWriter, AudioPredictiveCodingModel,       # noqa: F401
                            cuda0_writing_condition, load_to_cpu, num_parameters)
