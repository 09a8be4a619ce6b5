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


def test_cluster_kernel_workspaces_and_modes(lib):
    """Host logic of the small-batch cluster kernels that needs no GPU: the observe workspace holds the 16 per-rank weight
    images (136,192 bytes each at the default sizes), the backward workspace the 16 transposed images (128,000 bytes each)
    plus two floats per sequence; sizes outside the cluster geometry get the minimal backward workspace (the entry point
    then runs the per-sequence kernel); an unknown mode is refused before anything is launched."""
    from repo_b200._lib import Dims
    d = Dims(200, 30, 6, 200, 1024)
    assert lib.repo_b200_observe_workspace_bytes(C.byref(d), 49, 50) >= 16 * 136192
    need = lib.repo_b200_observe_bwd_workspace_bytes(C.byref(d), 50)
    assert need >= 16 * 128000 + 50 * 8 and need % 256 == 0
    odd = Dims(64, 8, 4, 256, 16)      # belief and hidden in different multiples of 16: no cluster geometry
    assert lib.repo_b200_observe_bwd_workspace_bytes(C.byref(odd), 50) == 256
    rc = lib.repo_b200_observe_bwd_ws(C.byref(d), None, *([None] * 24), 5, 3, 1, 1, 0.1, None, 0, 7, None)
    assert rc == -1 and b"mode" in lib.repo_b200_last_error()


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


def test_replay_buffer_host_path_matches_reference_fixture(golden_dir):
    """repo_b200.replay.SequenceReplayBuffer.sample consumes np.random and indexes exactly like
    common/buffers.py:156-166 (fixture generated by the reference class)."""
    import numpy as np
    from repo_b200.replay import SequenceReplayBuffer
    g = np.load(os.path.join(golden_dir, "replay_indices.npz"))
    for tag, seed in (("partial", 3), ("full", 4), ("fullwrap", 5)):
        cap, n_push, B, L, pos, full, length = [int(v) for v in g[f"{tag}_meta"]]
        buf = SequenceReplayBuffer(cap, (2,), (1,))
        for i in range(n_push):
            buf.push(np.array([i, -i], np.float32), np.array([i * 0.5], np.float32), float(i), float(i % 11 == 0))
        assert (buf.pos, buf.full, len(buf)) == (pos, bool(full), length)
        np.random.seed(seed)
        obs, act, rew, done = buf.sample(B, L)
        for name, got in (("obs", obs), ("act", act), ("rew", rew), ("done", done)):
            np.testing.assert_array_equal(got, g[f"{tag}_{name}"])


def test_replay_buffer_save_load_roundtrip(tmp_path):
    import numpy as np
    from repo_b200.replay import SequenceReplayBuffer
    buf = SequenceReplayBuffer(8, (3, 4, 4), (2,), obs_type=np.uint8)
    for i in range(11):
        buf.push(np.full((3, 4, 4), i, np.uint8), np.array([i, -i], np.float32), float(i), 0.0)
    buf.save(str(tmp_path / "buffer.npz"))
    keys = set(np.load(str(tmp_path / "buffer.npz")).files)
    assert keys == {"capacity", "observations", "actions", "rewards", "dones", "pos", "full"}  # buffers.py:193 format
    b2 = SequenceReplayBuffer(8, (3, 4, 4), (2,), obs_type=np.uint8)
    b2.load(str(tmp_path / "buffer.npz"))
    assert b2.pos == buf.pos and b2.full and b2.dones[b2.pos - 1] == 1
    np.testing.assert_array_equal(b2.observations, buf.observations)
