#!/usr/bin/env python
"""Determined M = K = 6 mixtures around the shared-memory limit of the resident loop (T = 116: two V buffers, tracked-inverse
sweep; T = 118, 124: one buffer or no residency -> the pair sweep / the kernel-per-step loop) against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import overiva_b200 as ob
from oracle import overiva_oracle as orc
from overiva_b200.synth import stft_domain_batch_torch

for T in (116, 118, 124):
    X = stft_domain_batch_torch(1, T, 2049, 6, 6, seed=T, device=torch.device("cuda", 0), chunk=1)[0].cpu().numpy()
    Yo, Wo = orc.overiva(X, n_iter=8, return_filters=True)
    Y, W = ob.overiva(X, n_iter=8, return_filters=True)
    eY, eW = np.abs(Y - Yo).max() / np.abs(Yo).max(), np.abs(W - Wo).max() / np.abs(Wo).max()
    print("T", T, "rel err Y %.2e W %.2e" % (eY, eW))
    assert eY < 1e-10 and eW < 1e-10
