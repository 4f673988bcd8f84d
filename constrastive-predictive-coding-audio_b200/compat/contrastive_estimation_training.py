"""Reference module name ``contrastive_estimation_training`` (contrastive_estimation_training.py:12-391) -> cpc_b200."""
import _bootstrap  # noqa: F401
from cpc_b200.trainer import (ContrastiveEstimationTrainer, DeterministicSampler,                    # noqa: F401
                              difference_score_function, linear_score_function, softplus_score_function)


def grad_mean_var(module):
    """contrastive_estimation_training.py:385-391."""
    import torch
    return {name: [torch.mean(p.grad).item(), torch.var(p.grad).item()]
            for name, p in module.named_parameters() if p.grad is not None}
