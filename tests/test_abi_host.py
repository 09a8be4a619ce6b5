"""CPU: the C-ABI library loads and exports every symbol include/repo_b200.h declares; host-side
shape logic and error behaviour that needs no GPU."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from repo_b200 import build, _lib
    build.build()  # nvcc cross-compiles without a GPU
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "repo_b200.h")).read()
    declared = set(re.findall(r"\b(repo_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/repo_b200.h but not exported"
    from repo_b200 import _lib
    assert declared == set(_lib.EXPORTS)


def test_workspace_sizes(lib):
    from repo_b200._lib import Dims
    d = Dims(200, 30, 6, 200, 1024)
    img = lib.repo_b200_imagine_workspace_bytes(C.byref(d))
    obs = lib.repo_b200_observe_workspace_bytes(C.byref(d), 49, 50)
    # 42 weight tiles / 540 k16 slabs of 8 KB + 38 bias tiles for the imagine program
    assert img >= 540 * 8192 and img % 256 == 0
    assert obs >= 49 * 50 * 200 * 4
    bad = Dims(300, 30, 6, 200, 1024)  # belief > 256 is outside the TMEM budget
    assert lib.repo_b200_imagine_workspace_bytes(C.byref(bad)) == 0
    assert b"belief_size" in lib.repo_b200_last_error()


def test_no_device_fails_loudly(lib):
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    assert lib.repo_b200_device_info(None, None, None) < 0
    assert b"CUDA" in lib.repo_b200_last_error()


def test_dims_recovered_from_state_dict():
    from oracle import rssm_oracle as O
    from repo_b200 import ops
    d = ops.dims_of(O.make_transition_params(0))
    assert (d.belief, d.state, d.action, d.hidden, d.embed) == (200, 30, 6, 200, 1024)


def test_module_interface_matches_reference_names():
    """state_dict keys / shapes of the drop-in TransitionModel."""
    from repo_b200.rssm import TransitionModel
    m = TransitionModel(200, 30, 6, 200, 1024, "elu")
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert sd == {
        "fc_embed_state_action.weight": (200, 36), "fc_embed_state_action.bias": (200,),
        "rnn.weight_ih": (600, 200), "rnn.weight_hh": (600, 200), "rnn.bias_ih": (600,), "rnn.bias_hh": (600,),
        "fc_embed_belief_prior.weight": (200, 200), "fc_embed_belief_prior.bias": (200,),
        "fc_state_prior.weight": (60, 200), "fc_state_prior.bias": (60,),
        "fc_embed_belief_posterior.weight": (200, 1224), "fc_embed_belief_posterior.bias": (200,),
        "fc_state_posterior.weight": (60, 200), "fc_state_posterior.bias": (60,),
    }
    assert sum(p.numel() for p in m.parameters()) == 557920
    for name in ("observe", "imagine", "compute_belief", "compute_prior_state", "compute_posterior_state", "obs_step", "img_step"):
        assert callable(getattr(m, name))
    with pytest.raises(RuntimeError):
        TransitionModel(200, 30, 6, 200, 1024, "gelu")  # unsupported activation: no silent fallback


def test_cpu_call_raises_not_falls_back():
    from oracle import rssm_oracle as O
    from repo_b200.rssm import TransitionModel
    m = TransitionModel(32, 8, 3, 24, 40, "elu")
    x = O.make_observe_inputs(0, 4, 2, dict(belief=32, state=8, action=3, hidden=24, embed=40))
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m.observe(x["prev_belief"], x["prev_state"], x["actions"], x["embeds"], x["nonterms"].squeeze(-1))
