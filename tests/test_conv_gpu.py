"""GPU: conv encoder / decoder (implicit GEMM on the layer machine) vs the reference fixture and the oracle
(torch conv ops on CPU), forward and backward.  Tolerance: rtol 1e-3 with an absolute floor of 1e-4 of the
tensor's max magnitude (activations shrink through a default-initialised conv stack)."""
import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from repo_b200 import synth
from tests import _cases as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rel_close(got, want, name, rtol=1e-3, floor=1e-4):
    got, want = got.detach().cpu().double().numpy(), np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max() + 1e-30
    np.testing.assert_allclose(got / scale, want / scale, rtol=rtol, atol=floor, err_msg=name)


def test_visual_encoder_forward_matches_reference(dev):
    from repo_b200.conv import VisualEncoder
    g, _ = C.load("conv_stacks")
    p = synth.make_conv_params("encoder", 700)
    enc = VisualEncoder(1024).to(dev)
    enc.load_state_dict(p)
    x = synth.make_frames(702, 3)
    with torch.no_grad():
        e = enc(x.to(dev))
    assert e.shape == (3, 1024)
    rel_close(e, g["embed"], "embed vs reference VisualEncoder")
    rel_close(e, O.visual_encoder(p, x), "embed vs oracle")
    # a batch that is not a multiple of any row tile, and more frames than one wave of CTAs needs
    x2 = synth.make_frames(705, 37)
    with torch.no_grad():
        rel_close(enc(x2.to(dev)), O.visual_encoder(p, x2), "embed 37 frames")


def test_visual_encoder_backward_matches_autograd(dev):
    from repo_b200.conv import VisualEncoder
    p = synth.make_conv_params("encoder", 710)
    x = synth.make_frames(711, 5)
    R = torch.from_numpy(np.random.RandomState(1).standard_normal((5, 1024)).astype(np.float32))
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    (O.visual_encoder(p64, x.double()) * R.double()).sum().backward()
    enc = VisualEncoder(1024).to(dev)
    enc.load_state_dict(p)
    (enc(x.to(dev)) * R.to(dev)).sum().backward()
    for k, w in p64.items():
        rel_close(dict(enc.named_parameters())[k].grad, w.grad, "encoder grad " + k, rtol=2e-3, floor=3e-4)
