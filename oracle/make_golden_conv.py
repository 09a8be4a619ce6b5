"""Golden fixtures for the conv stacks from the UNMODIFIED reference modules (build container only).
    python oracle/make_golden_conv.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import load_reference  # noqa: E402
from repo_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    torch.set_num_threads(1)
    R = load_reference()
    import importlib.util
    spec = importlib.util.spec_from_file_location("refmodels.encoder", "/root/reference/algorithms/repo/models/encoder.py")
    enc_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(enc_mod)
    enc = enc_mod.VisualEncoder(1024, "relu")
    enc.load_state_dict(synth.make_conv_params("encoder", 700))
    dec = R.decoder.VisualObservationModel(200, 30, 1024, "relu")
    dec.load_state_dict(synth.make_conv_params("decoder", 701))
    x = synth.make_frames(702, 3)
    b, s = synth.make_imagine_inputs(703, 3, 2)["belief"], synth.make_imagine_inputs(703, 3, 2)["state"]
    with torch.no_grad():
        e = enc(x)
        o = dec(b, s)
    np.savez_compressed(os.path.join(OUT, "conv_stacks.npz"), embed=e.numpy(), recon=o.numpy())
    print("embed", tuple(e.shape), float(e.abs().mean()), "recon", tuple(o.shape), float(o.abs().mean()))


if __name__ == "__main__":
    main()
