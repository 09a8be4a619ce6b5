"""Rebuild the seeded inputs that belong to a golden fixture (see oracle/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import rssm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

OBSERVE_CASES = ["observe_T8_B10", "observe_T8_B10_hot", "observe_default_tail", "observe_prior_only",
                 "observe_no_nonterm", "observe_tiny_dims", "observe_T2_B1"]
IMAGINE_CASES = ["imagine_N24_H6", "imagine_N8_H15", "imagine_N16_H15_hot", "imagine_tiny_dims"]
OBS_NAMES = ["beliefs", "prior_states", "prior_means", "prior_std_devs",
             "posterior_states", "posterior_means", "posterior_std_devs"]
IMG_NAMES = ["beliefs", "prior_states", "prior_means", "prior_std_devs"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    data = {k: z[k] for k in z.files if not k.startswith("meta_")}
    meta = {k[5:]: z[k].item() for k in z.files if k.startswith("meta_")}
    return data, meta


def dims_of(meta):
    return {k: int(meta["dim_" + k]) for k in ("belief", "state", "action", "hidden", "embed")}


def observe_case(name):
    data, meta = load(name)
    dims = dims_of(meta)
    seed = int(meta["seed"])
    params = O.make_transition_params(seed, dims, float(meta["scale"]))
    x = O.make_observe_inputs(seed + 10, int(meta["T"]), int(meta["B"]), dims,
                              p_done=float(meta["p_done"]), embed_scale=float(meta["embed_scale"]))
    if not meta["use_obs"]:
        x["embeds"] = None
        x["eps_post"] = None
    if not meta["use_nt"]:
        x["nonterms"] = None
    return params, x, data, meta


def imagine_case(name):
    data, meta = load(name)
    dims = dims_of(meta)
    seed = int(meta["seed"])
    D, S, A, H = dims["belief"], dims["state"], dims["action"], dims["hidden"]
    scale = float(meta["scale"])
    params = O.make_transition_params(seed, dims, scale)
    actor = O.make_mlp_params(seed + 1, D + S, H, 2 * A, 4, scale)
    reward = O.make_mlp_params(seed + 2, D + S, H, 1, 3, scale)
    value = O.make_mlp_params(seed + 3, D + S, H, 1, 3, scale)
    x = O.make_imagine_inputs(seed + 20, int(meta["N"]), int(meta["H"]), dims)
    return params, actor, reward, value, x, data, meta


def t(a):
    return torch.from_numpy(np.asarray(a))
