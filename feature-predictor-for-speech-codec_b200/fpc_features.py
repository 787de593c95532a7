"""What the reference's callers do right after `encoder` (generate_qtz_features.py:59-70,
synthesis_qtz.py:158-160): rescale the decoded features, derive the LPC coefficients from the
cepstrum and lay the result out for LPCNet -- here on the device.

    feats = lpcnet_features(c_in)             # (B, L, 36) = [20 features * MAXI | 16 LPC]
    chunks = lpcnet_chunks(feats)             # (B, n, 19, 36) windows of 19 frames hopping by 15

`lpcnet_chunks(feats[0], n_chunks=10)` is the reference's
`np.lib.stride_tricks.as_strided(all_features.flatten(), shape=(10, 19, 36), strides=(15*36*s, 36*s, s))`
for its single-utterance batches, as a strided VIEW (no copy) with the bounds checked.
"""
import torch

import fpc_native as N
from ceps2lpc.ceps2lpc_vct import ceps2lpc_device

MAXI = 24.1            # dataset normalisation constant (dataset_syn.py:27, synthesis_qtz.py:37)
CHUNK_FRAMES = 19
CHUNK_HOP = 15


def lpcnet_features(c_in, maxi=MAXI):
    """(B, L, 20) decoded, normalised features (CUDA) -> (B, L, 36): the features times `maxi`, then the 16 LPC
    coefficients of each frame (fpc_ceps2lpc)."""
    N.require_cuda()
    if not (isinstance(c_in, torch.Tensor) and c_in.is_cuda):
        raise N.FpcError("lpcnet_features needs a CUDA tensor: no CPU fallback")
    if c_in.dim() != 3 or c_in.shape[2] != 20:
        raise ValueError("c_in must be (batch, frames, 20), got %r" % (tuple(c_in.shape),))
    x = (c_in.detach().to(torch.float32) * float(maxi)).contiguous()      # feat_in *= MAXI  (generate_qtz_features.py:59)
    lpc, _, _ = ceps2lpc_device(x.reshape(-1, 20))
    return torch.cat((x, lpc.reshape(x.shape[0], x.shape[1], 16)), -1)


def lpcnet_chunks(features, n_chunks=None, frames=CHUNK_FRAMES, hop=CHUNK_HOP):
    """(L, 36) or (B, L, 36) -> (n, frames, 36) or (B, n, frames, 36) overlapping windows (a view).  n_chunks=None
    takes every complete window; asking for more windows than fit raises (the reference's as_strided would read past
    the end of the buffer)."""
    squeeze = features.dim() == 2
    f = features.unsqueeze(0) if squeeze else features
    if f.dim() != 3:
        raise ValueError("features must be (L, C) or (B, L, C)")
    L = f.shape[1]
    avail = (L - frames) // hop + 1 if L >= frames else 0
    n = avail if n_chunks is None else int(n_chunks)
    if n > avail:
        raise ValueError("%d frames hold %d windows of %d frames hopping by %d, %d requested" % (L, avail, frames, hop, n))
    f = f.contiguous()
    B, _, C = f.shape
    out = f.as_strided((B, n, frames, C), (L * C, hop * C, C, 1))
    return out[0] if squeeze else out
