"""Receiver side of the codec: wire format of the per-frame index record and the decoder that
consumes it.

The reference returns only histograms from `Wavernn.encoder`, defines no bitstream, and its
`decoder` / `dec_features` are not runnable (models/wavernn.py:367-379,
generate_qtz_features.py:74-91).  Here the encoder's index record `idx (B, L, 4) int32`
(`Wavernn.last_result.idx`) is the transmitted quantity:

    words = pack_frames(idx)                       # one uint32 per 10 ms frame = 3.2 kbit/s before entropy coding
    idx   = unpack_frames(cfg, words)              # the codebook set restores the "nothing coded" entries
    r_qtz = dequantize(cfg, idx)                   # == the encoder's own r_qtz, bit for bit
    c     = decode_indices(model, cfg, idx, pitch) # == the encoder's c_in, bit for bit

All of it runs on the device through the C ABI (fpc_pack_frames, fpc_unpack_frames,
fpc_dequantize, fpc_decode); there is no CPU fallback.
"""
import torch

import fpc_codebooks
import fpc_native as N

WORD_BITS = {"ind1": (0, 1), "ind2": (1, 1), "scalar": (2, 8), "vq1": (10, 10), "vq2": (20, 10)}


def _idx_on_device(idx):
    N.require_cuda()
    t = torch.as_tensor(idx)
    if not t.is_cuda:
        t = t.cuda()
    if t.shape[-1] != 4:
        raise ValueError("index record must be (..., 4), got %r" % (tuple(t.shape),))
    return t.to(torch.int32).contiguous()


def pack_frames(idx):
    """(..., 4) int32 index record -> (...) int32 tensor holding the 32-bit frame words."""
    t = _idx_on_device(idx)
    n = t.numel() // 4
    words = torch.empty(t.shape[:-1], dtype=torch.int32, device=t.device)
    with torch.cuda.device(t.device):
        N.check(N.lib().fpc_pack_frames(t.data_ptr(), n, words.data_ptr(), N.current_stream(t.device)), "fpc_pack_frames")
    return words


def unpack_frames(cfg, words, device=None):
    """Frame words -> (..., 4) int32 index record; `cfg` names the codebook files (as for Wavernn.encoder)."""
    N.require_cuda()
    w = torch.as_tensor(words)
    if not w.is_cuda:
        w = w.cuda() if device is None else w.to(device)
    w = w.to(torch.int32).contiguous()
    cbs = fpc_codebooks.from_cfg(cfg, w.device)
    idx = torch.empty(tuple(w.shape) + (4,), dtype=torch.int32, device=w.device)
    with torch.cuda.device(w.device):
        N.check(N.lib().fpc_unpack_frames(cbs.ptr(), w.data_ptr(), w.numel(), idx.data_ptr(), N.current_stream(w.device)),
                "fpc_unpack_frames")
        torch.cuda.current_stream(w.device).synchronize()      # cbs may be dropped by the caller
    return idx


def dequantize(cfg, idx):
    """Index record -> r_qtz (..., 18) float32: scalar table entry for c0, codeword sums for c1..c17, zeros where
    nothing was coded (models/wavernn.py:217-240 read backwards)."""
    t = _idx_on_device(idx)
    cbs = fpc_codebooks.from_cfg(cfg, t.device)
    out = torch.empty(tuple(t.shape[:-1]) + (18,), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        N.check(N.lib().fpc_dequantize(cbs.ptr(), t.data_ptr(), t.numel() // 4, out.data_ptr(), N.current_stream(t.device)),
                "fpc_dequantize")
        torch.cuda.current_stream(t.device).synchronize()
    return out


def decode_indices(model, cfg, idx, pitch):
    """The decoder: index record (B, L, 4) + pitch dims (B, L, 2) [or the (B, L, 20) features, of which only the
    last two dims are read] -> decoded features (B, L, 20).  Equals the encoder's c_in bit for bit."""
    r_qtz = dequantize(cfg, idx)
    p = torch.as_tensor(pitch).to(r_qtz.device)
    return model.decoder(cfg, p, r_qtz)


def bits_per_frame():
    """Fixed-length wire format: 30 payload bits in a 32-bit word (see fpc_bitrate for the entropy-coded bound)."""
    return 32
