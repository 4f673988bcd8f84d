"""Reference module name ``setup_functions``: ``setup_model`` (setup_functions.py:68-117) with the reference's
signature and defaults, returning (pc_model, preprocessing_module, untraced_model).  One process per GPU replaces
nn.DataParallel (cpc_b200.ddp), so no wrapping happens here."""
import _bootstrap  # noqa: F401
from audio_model import *                                                                            # noqa: F401,F403
from contrastive_estimation_training import *                                                        # noqa: F401,F403
from scalogram_model import *                                                                        # noqa: F401,F403
from cpc_b200 import configs as _configs


def setup_model(cqt_params=None, encoder_params=None, ar_params=None, trainer_args=None, device=None, visible_steps=60,
                prediction_steps=16, trace_model=False, use_all_GPUs=True, activation_register=None):
    if cqt_params is None or encoder_params is None or ar_params is None or trainer_args is None:
        # the reference's defaults are the dicts of its own configs package
        from configs.cqt_configs import cqt_default_dict
        from configs.scalogram_resnet_configs import scalogram_resnet_architecture_1
        from configs.autoregressive_model_configs import ar_conv_default_dict
        from configs.contrastive_estimation_configs import contrastive_estimation_default_dict
        cqt_params = cqt_default_dict if cqt_params is None else cqt_params
        encoder_params = scalogram_resnet_architecture_1 if encoder_params is None else encoder_params
        ar_params = ar_conv_default_dict if ar_params is None else ar_params
        trainer_args = contrastive_estimation_default_dict if trainer_args is None else trainer_args
    return _configs.setup_model(cqt_params, encoder_params, ar_params, trainer_args, device=device,
                                visible_steps=visible_steps, prediction_steps=prediction_steps, trace_model=trace_model,
                                use_all_GPUs=use_all_GPUs, activation_register=activation_register)
