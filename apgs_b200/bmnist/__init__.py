"""Bouncing-MNIST tracking task: drop-in for the hot-path functions of apgs/bmnist/{objectives,models,affine_transformer}.py."""
from .affine_transformer import Affine_Transformer  # noqa: F401
from .models import Enc_coor, Enc_digit, Dec_coor, Dec_digit, init_models, save_models  # noqa: F401
from .objectives import apg_objective, oneshot, apg_where, apg_what, propose_one_movement, resample_variables, bpg_objective, bpg_what  # noqa: F401
