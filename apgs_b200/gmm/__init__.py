"""GMM clustering task: drop-in for the hot-path functions of apgs/gmm/{objectives,models,kls_gmm}.py."""
from .models import Enc_rws_eta, Enc_apg_eta, Enc_apg_z, Generative, init_apg_models, save_apg_models  # noqa: F401
from .objectives import (apg_objective, apg_objective_streams, oneshot, apg_update_eta, apg_update_z, apg_update_joint,  # noqa: F401
                         resample_variables, gibbs_objective, gibbs_sweep, bpg_objective, bpg_update_eta, posterior_z)
