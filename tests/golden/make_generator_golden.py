#!/usr/bin/env python3
"""Golden vector of the reference's GAN ``Generator`` (paule/models.py:594-652), used in the prologue of plan_resynth.
Run in the build container only:  python tests/golden/make_generator_golden.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from paule import models as ref_models  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
for tag, osz, length in (("cp", 30, 46), ("mel", 60, 23)):
    torch.manual_seed(4)
    gen = ref_models.Generator(output_size=osz).eval()
    g = torch.Generator().manual_seed(9)
    noise, vec = torch.randn(2, 1, 100, generator=g), torch.randn(2, 300, generator=g)
    with torch.no_grad():
        out[f"{tag}_y"] = gen(noise, length, vec).numpy()
    out[f"{tag}_noise"], out[f"{tag}_vec"] = noise.numpy(), vec.numpy()
path = os.path.join(HERE, "generator_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path))
