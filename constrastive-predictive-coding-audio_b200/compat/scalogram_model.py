"""Reference module name ``scalogram_model`` (scalogram_model.py:11-544) -> cpc_b200."""
import _bootstrap  # noqa: F401
from constant_q_transform import *                                                                   # noqa: F401,F403
from audio_model import *                                                                            # noqa: F401,F403
from cpc_b200.frontend import PreprocessingModule                                                    # noqa: F401
from cpc_b200.encoders import (Conv2dSeparable, ScalogramEncoder, ScalogramEncoderBlock,             # noqa: F401
                               ScalogramResidualEncoder, cqt_default_dict, default_encoder_block_dict,
                               scalogram_encoder_default_dict)
