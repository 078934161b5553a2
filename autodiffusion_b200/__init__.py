"""B200-native (sm_100a) candidate evaluator for AutoDiffusion.

Drop-in for the hot path of lilijiangg/AutoDiffusion's guided_diffusion example:
`create_model_and_diffusion`, the diffusion object's `ddim_sample_loop`, the
`Dynamic_UNetModel` with its per-call block-skip list, and the search's candidate
evaluation. All compute runs in hand-written CUDA behind the C-ABI of
`include/adb200.h`; there is no CPU or PyTorch fallback.
"""
from .script_util import (  # noqa: F401
    NUM_CLASSES,
    add_dict_to_argparser,
    args_to_dict,
    classifier_and_diffusion_defaults,
    classifier_defaults,
    create_classifier,
    create_classifier_and_diffusion,
    create_gaussian_diffusion,
    create_model,
    create_model_and_diffusion,
    diffusion_defaults,
    model_and_diffusion_defaults,
    str2bool,
)

__version__ = "0.1.0"
