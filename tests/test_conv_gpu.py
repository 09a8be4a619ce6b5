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


def test_visual_decoder_forward_matches_reference(dev):
    from repo_b200.conv import VisualObservationModel
    g, _ = C.load("conv_stacks")
    p = synth.make_conv_params("decoder", 701)
    dec = VisualObservationModel(200, 30, 1024).to(dev)
    dec.load_state_dict(p)
    xi = synth.make_imagine_inputs(703, 3, 2)
    with torch.no_grad():
        o = dec(xi["belief"].to(dev), xi["state"].to(dev))
    assert o.shape == (3, 3, 64, 64)
    rel_close(o, g["recon"], "recon vs reference VisualObservationModel")
    rel_close(o, O.visual_decoder(p, xi["belief"], xi["state"]), "recon vs oracle")


def test_visual_decoder_backward_matches_autograd(dev):
    from repo_b200.conv import VisualObservationModel
    p = synth.make_conv_params("decoder", 720)
    xi = synth.make_imagine_inputs(721, 4, 2)
    R = torch.from_numpy(np.random.RandomState(2).standard_normal((4, 3, 64, 64)).astype(np.float32))
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    b64, s64 = xi["belief"].double().requires_grad_(True), xi["state"].double().requires_grad_(True)
    (O.visual_decoder(p64, b64, s64) * R.double()).sum().backward()
    dec = VisualObservationModel(200, 30, 1024).to(dev)
    dec.load_state_dict(p)
    gb, gs = xi["belief"].to(dev).requires_grad_(True), xi["state"].to(dev).requires_grad_(True)
    (dec(gb, gs) * R.to(dev)).sum().backward()
    for k, w in p64.items():
        rel_close(dict(dec.named_parameters())[k].grad, w.grad, "decoder grad " + k, rtol=2e-3, floor=3e-4)
    rel_close(gb.grad, b64.grad, "d belief", rtol=2e-3, floor=3e-4)
    rel_close(gs.grad, s64.grad, "d state", rtol=2e-3, floor=3e-4)


@pytest.mark.parametrize("case", ["enc1_nchw_c3", "enc2_stride2", "enc4_k2048_n256", "dec2_k1152_n256", "dec3_k576_n128", "dec4_n12_ld16"])
def test_conv_wgrad_matches_materialised_gemm(dev, case):
    """tcgen05 weight-gradient kernel (MN-major operands, atomic row-slice reduction) against im2col + fp64 GEMM, with
    gradient magnitudes (1e-6) far below fp16's normal range to exercise the power-of-two operand scaling."""
    from repo_b200 import conv as cv
    Fr = 5
    enc = cv._enc_maps((64, 64))
    shapes = {
        "enc1_nchw_c3": ((Fr, 3, 64, 64), enc[0], 32, None),
        "enc2_stride2": ((Fr, 31, 31, 32), enc[1], 64, None),
        "enc4_k2048_n256": ((Fr, 6, 6, 128), enc[3], 256, None),
        "dec2_k1152_n256": ((Fr, 5, 5, 128), cv._deconv_map(128, 5, 5, 5, False, True), 256, None),
        "dec3_k576_n128": ((Fr, 13, 13, 64), cv._deconv_map(64, 13, 13, 6, False, True), 128, None),
        "dec4_n12_ld16": ((Fr, 30, 30, 32), cv._deconv_map(32, 30, 30, 6, True, False), 12, 16),
    }
    xs, cm, n_total, ld = shapes[case]
    rs = np.random.RandomState(11)
    x = torch.from_numpy(rs.standard_normal(xs).astype(np.float32)).to(dev)
    rows = Fr * cm.RA * cm.RB
    g = torch.zeros(rows, ld or n_total, device=dev)
    g[:, :n_total] = torch.from_numpy((rs.standard_normal((rows, n_total)) * 1e-6).astype(np.float32)).to(dev)
    want = g[:, :n_total].double().t() @ cv.im2col(x, Fr, cm).double()
    got = cv.conv_wgrad(x, g, Fr, n_total, cm)
    assert got.shape == (n_total, cm.K)
    rel_close(got, want.cpu().numpy(), case, rtol=1e-3, floor=1e-5)


def test_split_activation_format_is_equivalent_to_fp32(dev):
    """HL (fp16 hi/lo plane) operands: gathering an HL input, writing an HL output and masking with an HL activation must
    reproduce the fp32 paths (bit-exact for the gather / mask, 1 ulp for the re-joined output), and tiny positive
    activations must survive in the hi plane (it doubles as the ReLU mask of the backward pass)."""
    from repo_b200 import conv as cv
    torch.manual_seed(0)
    F_ = 7

    def to_hl(x):
        hi = x.half()
        return torch.stack([hi, (x - hi.float()).half()]).contiguous()

    cm3 = cv._deconv_map(64, 13, 13, 6, False, True)
    a2 = torch.relu(torch.randn(F_, 13, 13, 64, device=dev))
    w3, b3 = torch.randn(128, cm3.K, device=dev) * 0.05, torch.randn(128, device=dev) * 0.1
    y_f, y_hl_in = torch.empty(F_, 30, 30, 32, device=dev), torch.empty(F_, 30, 30, 32, device=dev)
    y_hl_out = cv.hl_empty((F_, 30, 30, 32), dev)
    cv.conv_gemm(a2, w3, b3, y_f, F_, 128, cm3)
    cv.conv_gemm(to_hl(a2), w3, b3, y_hl_in, F_, 128, cm3)
    cv.conv_gemm(to_hl(a2), w3, b3, y_hl_out, F_, 128, cm3)
    assert torch.equal(y_f, y_hl_in)
    assert (cv.hl_to_float(y_hl_out) - y_f).abs().max().item() <= 2.4e-7 * max(1.0, y_f.abs().max().item())
    assert torch.equal(y_hl_out[0] > 0, y_f > 0)          # mask semantics preserved, incl. values below fp16's range
    # tiny positive outputs: bias 1e-9 on zero weights -> relu output 1e-9 must still read as "positive"
    tiny = cv.hl_empty((F_, 30, 30, 32), dev)
    cv.conv_gemm(a2, torch.zeros_like(w3), torch.full((128,), 1e-9, device=dev), tiny, F_, 128, cm3)
    assert bool((tiny[0] > 0).all())
    # data gradient with an fp32 vs HL ReLU mask
    cm4 = cv._deconv_map(32, 30, 30, 6, True, False)
    G = torch.randn(F_, cm4.RA, cm4.RB, 16, device=dev) * 1e-4
    wd = torch.randn(32, 9 * 16, device=dev) * 0.05
    a3 = torch.relu(torch.randn(F_, 30, 30, 32, device=dev))
    dmap = cv.ConvMap(RA=30, RB=30, in_nchw=0, C=16, H=cm4.RA, W=cm4.RB, TH=3, TW=3, sy=1, sx=1, dy=1, dx=1, Ho=30, Wo=30)
    sc = cv._as_input_side(cv.grad_scales(G))
    o1, o2 = torch.empty(F_, 30, 30, 32, device=dev), torch.empty(F_, 30, 30, 32, device=dev)
    cv.conv_gemm(G, wd, None, o1, F_, 32, dmap, relu_mask=a3, scales=sc)
    cv.conv_gemm(G, wd, None, o2, F_, 32, dmap, relu_mask=to_hl(a3), scales=sc)
    assert torch.equal(o1, o2)
    # weight gradient with an fp32 vs HL input
    g = torch.randn(F_ * cm3.RA * cm3.RB, 128, device=dev) * 1e-5
    d1, d2 = cv.conv_wgrad(a2, g, F_, 128, cm3), cv.conv_wgrad(to_hl(a2), g, F_, 128, cm3)
    assert ((d1 - d2).abs().max() / d1.abs().max()).item() < 2e-6   # atomics: summation order differs run to run


def test_decoder_backward_large_batch_branch_matches_small_batch_branch(dev):
    """From 256 frames the decoder's three plain-GEMM gradients (first transposed conv and fc1) run on the tcgen05
    kernels instead of cuBLAS: the gradients of one 256-frame batch must equal the summed gradients of its four 64-frame
    quarters (the small-batch branch, itself checked against fp64 autograd above)."""
    from repo_b200.conv import VisualObservationModel
    p = synth.make_conv_params("decoder", 730)
    xi = synth.make_imagine_inputs(731, 256, 2)
    R = torch.from_numpy(np.random.RandomState(3).standard_normal((256, 3, 64, 64)).astype(np.float32)).to(dev) * 1e-3

    def run(sl):
        dec = VisualObservationModel(200, 30, 1024).to(dev)
        dec.load_state_dict(p)
        b, s = xi["belief"][sl].to(dev).requires_grad_(True), xi["state"][sl].to(dev).requires_grad_(True)
        (dec(b, s) * R[sl]).sum().backward()
        return {k: v.grad for k, v in dec.named_parameters()}, b.grad, s.grad

    big, gb, gs = run(slice(0, 256))
    parts = [run(slice(i, i + 64)) for i in range(0, 256, 64)]
    for k in big:
        rel_close(big[k], sum(pt[0][k] for pt in parts).cpu().numpy(), "decoder grad " + k, rtol=1e-3, floor=2e-4)
    rel_close(gb, torch.cat([pt[1] for pt in parts]).cpu().numpy(), "d belief", rtol=1e-3, floor=2e-4)
    rel_close(gs, torch.cat([pt[2] for pt in parts]).cpu().numpy(), "d state", rtol=1e-3, floor=2e-4)


def test_non_default_embedding_size(dev):
    """embedding_size != 1024: the encoder gains Linear(1024, E) (encoder.py:30,40) and the decoder's first two layers take E
    channels (decoder.py:34-35); forward vs the oracle, encoder gradients vs fp64 autograd."""
    from repo_b200.conv import VisualEncoder, VisualObservationModel
    E = 512
    pe, pd = synth.make_conv_params("encoder", 740, embedding_size=E), synth.make_conv_params("decoder", 741, embedding_size=E)
    enc, dec = VisualEncoder(E).to(dev), VisualObservationModel(200, 30, E).to(dev)
    enc.load_state_dict(pe)
    dec.load_state_dict(pd)
    x = synth.make_frames(742, 5)
    xi = synth.make_imagine_inputs(743, 5, 2)
    with torch.no_grad():
        rel_close(enc(x.to(dev)), O.visual_encoder(pe, x), "embedding (E=512)")
        rel_close(dec(xi["belief"].to(dev), xi["state"].to(dev)), O.visual_decoder(pd, xi["belief"], xi["state"]), "recon (E=512)")
    R = torch.from_numpy(np.random.RandomState(5).standard_normal((5, E)).astype(np.float32)) * 1e-4
    p64 = {k: v.double().requires_grad_(True) for k, v in pe.items()}
    (O.visual_encoder(p64, x.double()) * R.double()).sum().backward()
    (enc(x.to(dev)) * R.to(dev)).sum().backward()
    for k, w in p64.items():
        rel_close(dict(enc.named_parameters())[k].grad, w.grad, "encoder grad " + k, rtol=2e-3, floor=3e-4)
