"""DMM ring-mixture task: drop-in for the hot-path functions of apgs/dmm/{objectives,models}.py."""
from .models import Enc_rws_mu, Enc_apg_local, Enc_apg_mu, Decoder, init_apg_models, save_apg_models  # noqa: F401
from .objectives import apg_objective, oneshot, apg_update_mu, apg_update_local, resample_variables, bpg_objective, bpg_update_mu  # noqa: F401
