"""ctypes binding of librepo_b200.so (C-ABI in include/repo_b200.h).

There is no fallback: if the library is missing or a call fails, this raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# REPO_B200_PROFILING=1 (scripts/stage_clock.py only) loads the -DRB_STAGE_CLOCK build with in-kernel clock stamps
LIB_PATH = os.path.join(HERE, "librepo_b200_prof.so" if os.environ.get("REPO_B200_PROFILING") == "1" else "librepo_b200.so")

ACT_KINDS = {"relu": 0, "elu": 1}
WEIGHTS_PACKED = 1


class Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("belief", "state", "action", "hidden", "embed")]


_RSSM_FIELDS = [
    "fc_embed_state_action_w", "fc_embed_state_action_b",
    "rnn_w_ih", "rnn_w_hh", "rnn_b_ih", "rnn_b_hh",
    "fc_embed_belief_prior_w", "fc_embed_belief_prior_b",
    "fc_state_prior_w", "fc_state_prior_b",
    "fc_embed_belief_posterior_w", "fc_embed_belief_posterior_b",
    "fc_state_posterior_w", "fc_state_posterior_b",
]


class RssmWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _RSSM_FIELDS]


class MlpWeights(C.Structure):
    _fields_ = [("w", C.c_void_p * 5), ("b", C.c_void_p * 5), ("n_layers", C.c_int)]


EXPORTS = [
    "repo_b200_version", "repo_b200_last_error", "repo_b200_device_info", "repo_b200_debug_flags", "repo_b200_debug_clock",
    "repo_b200_imagine_workspace_bytes", "repo_b200_imagine_fwd", "repo_b200_imagine_stash_floats", "repo_b200_imagine_bwd", "repo_b200_imagine_cond_fwd", "repo_b200_imagine_cond_bwd",
    "repo_b200_observe_workspace_bytes", "repo_b200_observe_fwd", "repo_b200_observe_stash_floats", "repo_b200_observe_bwd",
    "repo_b200_observe_bwd_workspace_bytes", "repo_b200_observe_bwd_ws",
    "repo_b200_linear_workspace_bytes", "repo_b200_linear_fwd",
    "repo_b200_head_workspace_bytes", "repo_b200_head_fwd",
    "repo_b200_tanh_normal_entropy_fwd", "repo_b200_replay_gather",
    "repo_b200_cell_workspace_bytes", "repo_b200_cell_fwd",
    "repo_b200_sqnorm_accumulate", "repo_b200_colsum", "repo_b200_adam_clip_step", "repo_b200_adam_clip_step_dev", "repo_b200_conv_workspace_bytes", "repo_b200_conv_gemm", "repo_b200_conv_wgrad", "repo_b200_pow2_scale", "repo_b200_grad_unshuffle", "repo_b200_tia_mix_fwd", "repo_b200_tia_mix_bwd", "repo_b200_im2col",
    "repo_b200_mlp_workspace_bytes", "repo_b200_mlp_fwd", "repo_b200_mlp_bwd", "repo_b200_tanh_normal_entropy_bwd",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (python -m repo_b200.build). "
            "repo_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.repo_b200_version.restype = ci
    L.repo_b200_last_error.restype = C.c_char_p
    L.repo_b200_device_info.argtypes = [C.POINTER(ci)] * 3
    L.repo_b200_device_info.restype = ci
    L.repo_b200_debug_flags.argtypes = [ci]
    L.repo_b200_debug_flags.restype = None
    L.repo_b200_debug_clock.argtypes = [vp]
    L.repo_b200_debug_clock.restype = None
    L.repo_b200_imagine_workspace_bytes.argtypes = [C.POINTER(Dims)]
    L.repo_b200_imagine_workspace_bytes.restype = sz
    L.repo_b200_imagine_fwd.argtypes = (
        [C.POINTER(Dims), C.POINTER(RssmWeights), C.POINTER(MlpWeights), C.POINTER(MlpWeights), C.POINTER(MlpWeights)]
        + [vp] * 12 + [ci, ci, ci] + [cf] * 6 + [vp, vp, sz, ci, ci, vp])
    L.repo_b200_imagine_fwd.restype = ci
    L.repo_b200_imagine_cond_fwd.argtypes = (
        [C.POINTER(Dims), C.POINTER(RssmWeights), C.POINTER(MlpWeights), C.POINTER(MlpWeights), C.POINTER(MlpWeights)]
        + [vp, vp, vp, ci] + [vp] * 10 + [ci, ci, ci] + [cf] * 6 + [vp, vp, sz, ci, ci, vp])
    L.repo_b200_imagine_cond_fwd.restype = ci
    L.repo_b200_imagine_cond_bwd.argtypes = ([C.POINTER(Dims), C.POINTER(RssmWeights), C.POINTER(MlpWeights), ci] + [vp] * 23
                                             + [ci, ci, ci, cf, cf, cf, vp])
    L.repo_b200_imagine_cond_bwd.restype = ci
    L.repo_b200_imagine_stash_floats.argtypes = [C.POINTER(Dims)]
    L.repo_b200_imagine_stash_floats.restype = ci
    L.repo_b200_imagine_bwd.argtypes = ([C.POINTER(Dims), C.POINTER(RssmWeights), C.POINTER(MlpWeights)] + [vp] * 23
                                        + [ci, ci, ci, cf, cf, cf, vp])
    L.repo_b200_imagine_bwd.restype = ci
    L.repo_b200_observe_workspace_bytes.argtypes = [C.POINTER(Dims), ci, ci]
    L.repo_b200_observe_workspace_bytes.restype = sz
    L.repo_b200_observe_fwd.argtypes = (
        [C.POINTER(Dims), C.POINTER(RssmWeights)] + [vp] * 16 + [ci, ci, ci, cf, vp, sz, ci, ci, vp])
    L.repo_b200_observe_fwd.restype = ci
    L.repo_b200_observe_stash_floats.argtypes = [C.POINTER(Dims)]
    L.repo_b200_observe_stash_floats.restype = ci
    L.repo_b200_observe_bwd.argtypes = [C.POINTER(Dims), C.POINTER(RssmWeights)] + [vp] * 24 + [ci, ci, ci, ci, cf, vp]
    L.repo_b200_observe_bwd.restype = ci
    L.repo_b200_observe_bwd_workspace_bytes.argtypes = [C.POINTER(Dims), ci]
    L.repo_b200_observe_bwd_workspace_bytes.restype = sz
    L.repo_b200_observe_bwd_ws.argtypes = ([C.POINTER(Dims), C.POINTER(RssmWeights)] + [vp] * 24 + [ci, ci, ci, ci, cf, vp, sz, ci, vp])
    L.repo_b200_observe_bwd_ws.restype = ci
    L.repo_b200_head_workspace_bytes.argtypes = [C.POINTER(Dims)]
    L.repo_b200_head_workspace_bytes.restype = sz
    L.repo_b200_head_fwd.argtypes = [C.POINTER(Dims), C.POINTER(MlpWeights), vp, vp, vp, ci, ci, vp, sz, ci, ci, vp]
    L.repo_b200_head_fwd.restype = ci
    L.repo_b200_tanh_normal_entropy_fwd.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp]
    L.repo_b200_tanh_normal_entropy_fwd.restype = ci
    ll = C.c_longlong
    L.repo_b200_replay_gather.argtypes = [vp, vp, vp, vp, vp, ci, ci, ll, ci, ll, ci, ci, vp, vp, vp, vp, vp, vp]
    L.repo_b200_replay_gather.restype = ci
    L.repo_b200_cell_workspace_bytes.argtypes = [C.POINTER(Dims), ci]
    L.repo_b200_cell_workspace_bytes.restype = sz
    L.repo_b200_cell_fwd.argtypes = [C.POINTER(Dims), C.POINTER(RssmWeights)] + [vp] * 6 + [ci, ci, cf, vp, sz, vp]
    L.repo_b200_cell_fwd.restype = ci
    L.repo_b200_mlp_workspace_bytes.argtypes = [C.POINTER(Dims), ci, ci]
    L.repo_b200_mlp_workspace_bytes.restype = sz
    L.repo_b200_mlp_fwd.argtypes = [C.POINTER(Dims), C.POINTER(MlpWeights), vp, vp, vp, ci, vp, ci, ci, vp, sz, vp]
    L.repo_b200_mlp_fwd.restype = ci
    L.repo_b200_mlp_bwd.argtypes = [C.POINTER(Dims), C.POINTER(MlpWeights), vp, vp, ci, vp, vp, vp, vp, vp, ci, ci, vp]
    L.repo_b200_mlp_bwd.restype = ci
    L.repo_b200_tanh_normal_entropy_bwd.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]
    L.repo_b200_tanh_normal_entropy_bwd.restype = ci
    L.repo_b200_sqnorm_accumulate.argtypes = [vp, C.c_longlong, vp, vp]
    L.repo_b200_sqnorm_accumulate.restype = ci
    L.repo_b200_colsum.argtypes = [vp, C.c_longlong, ci, C.c_longlong, vp, vp]
    L.repo_b200_colsum.restype = ci
    L.repo_b200_adam_clip_step.argtypes = [vp, vp, vp, vp, C.c_longlong, vp, cf, cf, cf, cf, cf, ci, vp]
    L.repo_b200_adam_clip_step.restype = ci
    L.repo_b200_adam_clip_step_dev.argtypes = [vp, vp, vp, vp, C.c_longlong, vp, cf, cf, cf, cf, cf, vp, vp]
    L.repo_b200_adam_clip_step_dev.restype = ci
    L.repo_b200_conv_workspace_bytes.argtypes = [ci, ci]
    L.repo_b200_conv_workspace_bytes.restype = sz
    L.repo_b200_conv_gemm.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, C.POINTER(ci), ci, C.POINTER(ci), vp, sz, vp]
    L.repo_b200_conv_gemm.restype = ci
    L.repo_b200_conv_wgrad.argtypes = [vp, vp, vp, vp, ci, ci, ci, C.POINTER(ci), ci, vp]
    L.repo_b200_conv_wgrad.restype = ci
    L.repo_b200_pow2_scale.argtypes = [vp, C.c_longlong, C.c_float, ci, vp, vp, vp]
    L.repo_b200_pow2_scale.restype = ci
    L.repo_b200_tia_mix_fwd.argtypes = [vp, vp, vp, vp, vp, vp, C.c_longlong, ci, vp]
    L.repo_b200_tia_mix_fwd.restype = ci
    L.repo_b200_tia_mix_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_longlong, ci, vp]
    L.repo_b200_tia_mix_bwd.restype = ci
    L.repo_b200_grad_unshuffle.argtypes = [vp, ci, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
    L.repo_b200_grad_unshuffle.restype = ci
    L.repo_b200_im2col.argtypes = [vp, vp, ci, C.POINTER(ci), vp]
    L.repo_b200_im2col.restype = ci
    L.repo_b200_linear_workspace_bytes.argtypes = [ci, ci]
    L.repo_b200_linear_workspace_bytes.restype = sz
    L.repo_b200_linear_fwd.argtypes = [vp, ci, ci, ci, vp, vp, ci, vp, ci, vp, sz, ci, vp]
    L.repo_b200_linear_fwd.restype = ci
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().repo_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
