"""Drop-in `Wavernn` for the closed-loop predictive-coding path.

Mirrors /root/reference/src/models/wavernn.py: same constructor (:24), same sub-module names
and therefore the same state_dict keys (`rnn1.*_l0`, `rnn2.*_l0`, `dual_fc.0.*`, :37-38,48-52),
same `forward` (:63-102) and the same `encoder` signature and 7-tuple (:165-256).  The
difference is below the seam: `encoder` does not loop over frames and utterances in Python with
host-side NumPy quantisers; it makes ONE call into the C ABI (`fpc_encode`, include/fpc_b200.h),
whose persistent sm_100a kernel runs predictor, thresholds, scalar + m-best vector quantisers
and the feedback for every frame of every utterance.

There is no CPU fallback: `encoder` / `decoder` raise if `feat` is not a CUDA tensor.
`forward` is kept as plain torch modules for teacher-forced training (train_frame.py), which
is outside the hot path.
"""
import ctypes

import numpy as np
import torch
from torch import Tensor, nn

import fpc_codebooks
import fpc_native as N

device = 'cuda'   # the reference's module global (wavernn.py:20); kept for callers that set it

_WEIGHT_FIELDS = (
    ("w_ih1", "rnn1.weight_ih_l0"), ("w_hh1", "rnn1.weight_hh_l0"),
    ("b_ih1", "rnn1.bias_ih_l0"), ("b_hh1", "rnn1.bias_hh_l0"),
    ("w_ih2", "rnn2.weight_ih_l0"), ("w_hh2", "rnn2.weight_hh_l0"),
    ("b_ih2", "rnn2.bias_ih_l0"), ("b_hh2", "rnn2.bias_hh_l0"),
    ("w_fc", "dual_fc.0.weight"), ("b_fc", "dual_fc.0.bias"),
)
_GEOMETRY = (20, 384, 128, 18)   # FPC_IN_FEATURES, FPC_GRU1, FPC_GRU2, FPC_FC of this build


class EncodeResult:
    """Everything one fpc_encode call produced, still on the device."""
    __slots__ = ("c_in", "r", "r_qtz", "r_under", "ind1", "ind2", "idx", "codebooks")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


class Wavernn(nn.Module):

    def __init__(self, in_features=20, gru_units1=384, gru_units2=16, fc_units=20, attn_units=20, rnn_layers=2,
                 bidirectional=False, packing=False):
        super().__init__()
        self.scale = 1
        self.relu = nn.ReLU()
        self.packing = packing
        self.bidirectional = bidirectional
        self.rnn1 = nn.GRU(in_features, gru_units1, 1, bidirectional=bidirectional, batch_first=True)
        self.rnn2 = nn.GRU(gru_units1, gru_units2, 1, bidirectional=bidirectional, batch_first=True)
        self.dual_fc = nn.Sequential(nn.Linear(gru_units2, fc_units), nn.Tanh())
        self._geometry = (in_features, gru_units1, gru_units2, fc_units)
        self._packed = {}          # (device, precision) -> (version key, packed image tensor)
        self.precision = N.FPC_PREC_FP32
        self.last_result = None    # index record of the most recent encoder() call (EncodeResult with idx + codebooks only)

    # ------------------------------------------------------------------ reference: wavernn.py:63-102
    def forward(self, x: Tensor, h1=None, h2=None):
        x, h1 = self.rnn1(x, h1)
        x, h2 = self.rnn2(x, h2)
        if self.packing:
            from torch.nn.utils.rnn import pad_packed_sequence
            x, _ = pad_packed_sequence(x, batch_first=True)
        x = self.relu(x)
        x = torch.cat((x.unsqueeze(1), x.unsqueeze(1)), 1)
        x = self.dual_fc(x)
        x = torch.sum(x, dim=1)
        return x, h1, h2

    # ------------------------------------------------------------------ packed weights
    def _check_geometry(self):
        if self._geometry != _GEOMETRY or self.bidirectional:
            raise ValueError(
                "the fused closed-loop path is built for Wavernn(in_features=20, gru_units1=384, gru_units2=128, "
                "fc_units=18), unidirectional (synthesis_qtz.py:79-85); got %r bidirectional=%r"
                % (self._geometry, self.bidirectional))

    def packed_weights(self, dev, precision=None):
        """Device image of the parameters in the kernel's streaming order; rebuilt when any
        parameter changed (load_state_dict, optimiser step, .to())."""
        precision = self.precision if precision is None else precision
        self._check_geometry()
        sd = dict(self.named_parameters())
        vkey = tuple((sd[k].data_ptr(), sd[k]._version) for _, k in _WEIGHT_FIELDS)
        hit = self._packed.get((str(dev), precision))
        if hit is not None and hit[0] == vkey:
            return hit[1]
        L = N.lib()
        w = N.Weights()
        keep = []
        for field, key in _WEIGHT_FIELDS:
            t = sd[key].detach().to(device=dev, dtype=torch.float32).contiguous()
            keep.append(t)
            setattr(w, field, t.data_ptr())
        nbytes = L.fpc_packed_weights_bytes(precision)
        if nbytes == 0:
            raise N.FpcError("precision %d is not built into libfpc_b200" % precision)
        image = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            N.check(L.fpc_pack_weights(ctypes.byref(w), precision, image.data_ptr(), nbytes, N.current_stream(dev)),
                    "fpc_pack_weights")
            torch.cuda.current_stream(dev).synchronize()
        self._packed[(str(dev), precision)] = (vkey, image)
        return image

    # ------------------------------------------------------------------ reference: wavernn.py:165-256
    def encode_device(self, cfg, feat, mask, l1, l2, qtz=True, want_under=True, out=None):
        """The fused closed loop on device tensors.  Returns an EncodeResult (no host sync)."""
        N.require_cuda()
        if not (isinstance(feat, torch.Tensor) and feat.is_cuda):
            raise N.FpcError("Wavernn.encoder needs a CUDA tensor: the closed-loop path has no CPU fallback")
        if feat.dim() != 3 or feat.shape[2] != _GEOMETRY[0]:
            raise ValueError("feat must be (batch, frames, 20), got %r" % (tuple(feat.shape),))
        dev = feat.device
        feat = feat.detach().to(torch.float32).contiguous()
        B, Lf, _ = feat.shape
        cbs = fpc_codebooks.from_cfg(cfg, dev) if qtz else None
        if qtz and (cbs.arrays["vq"] is None or cbs.arrays["scl"] is None):
            # the reference would np.load('') here (wavernn.py:219,230)
            raise FileNotFoundError("cfg['cb_path'] and cfg['scl_cb_path'] are required when qtz is set")
        weights = self.packed_weights(dev)
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).detach().to(device=dev, dtype=torch.float32)
            if m.dim() == 2:           # (L, 2): the reference documents "(seq_length)" masks
                m = m.unsqueeze(0).expand(B, -1, -1)
            m = m.contiguous()
            if tuple(m.shape) != (B, Lf, 2):
                raise ValueError("mask must be (batch, frames, 2), got %r" % (tuple(m.shape),))
        o = out or {}

        def buf(name, shape, dtype=torch.float32):
            t = o.get(name)
            if t is None:
                t = torch.empty(shape, dtype=dtype, device=dev)
            return t

        res = EncodeResult(
            c_in=buf("c_in", (B, Lf, 20)), r=buf("r", (B, Lf, 18)), r_qtz=buf("r_qtz", (B, Lf, 18)),
            r_under=buf("r_under", (B, Lf, 18)) if want_under else None,
            ind1=buf("ind1", (B, Lf, 1)), ind2=buf("ind2", (B, Lf, 1)),
            idx=buf("idx", (B, Lf, 4), torch.int32), codebooks=cbs)
        io = N.EncodeIO()
        io.d_feat = feat.data_ptr()
        io.d_mask = m.data_ptr() if m is not None else None
        io.B, io.L = B, Lf
        io.l1, io.l2 = float(l1), float(l2)
        io.qtz = 1 if qtz else 0
        io.d_c_in, io.d_r, io.d_r_qtz = res.c_in.data_ptr(), res.r.data_ptr(), res.r_qtz.data_ptr()
        io.d_r_under = res.r_under.data_ptr() if res.r_under is not None else None
        io.d_ind1, io.d_ind2, io.d_idx = res.ind1.data_ptr(), res.ind2.data_ptr(), res.idx.data_ptr()
        with torch.cuda.device(dev):
            N.check(N.lib().fpc_encode(weights.data_ptr(), cbs.ptr() if cbs is not None else None, ctypes.byref(io),
                                       self.precision, None, 0, N.current_stream(dev)), "fpc_encode")
        # feat / m / weights must outlive the asynchronous kernel: tie them to the result
        res.codebooks = (cbs, feat, m, weights)
        # only the index record stays pinned by the module (the reference's 7-tuple has no indices; fpc_bitstream needs
        # them) -- not the ~1 GB of float outputs of a large call
        self.last_result = EncodeResult(idx=res.idx, codebooks=(cbs,))
        return res

    def encode_host(self, cfg, feat, l1, l2, qtz=True, out=None, chunks=0, device=None, want_under=False, want_hist=True):
        """Host-buffer form of the closed loop: what `feat.to('cuda')` -> `encoder(...)` -> `.cpu()` of the results
        does in the reference scripts (synthesis_qtz.py:149-160, generate_qtz_features.py:55-70), as ONE call that
        cuts the utterances along time and overlaps upload, kernel and download (C ABI fpc_encode_host).

        feat: (B, L, 20) float32 CPU tensor (pinned memory lets the copies overlap).  out: optional dict of CPU
        tensors to fill ("c_in", "r", "r_qtz", "r_under", "ind1", "ind2", "idx"); missing ones are allocated pinned
        ("r_under" only if want_under).  With want_hist the dict also gets "hist" (the raw usage counters, pinned int64)
        and "codebooks"; `host_cb_tot(res)` turns them into the reference's cb_tot list once the stream is synchronised.
        Returns the dict; the tensors are complete after torch.cuda.current_stream(device).synchronize().
        Bit-identical to encoder() on the same input."""
        N.require_cuda()
        if not isinstance(feat, torch.Tensor) or feat.is_cuda:
            raise N.FpcError("encode_host takes a CPU tensor; use encoder()/encode_device() for CUDA tensors")
        if feat.dim() != 3 or feat.shape[2] != _GEOMETRY[0]:
            raise ValueError("feat must be (batch, frames, 20), got %r" % (tuple(feat.shape),))
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        if dev.type != "cuda":
            raise N.FpcError("the model must live on a CUDA device (no CPU fallback)")
        feat = feat.detach().to(torch.float32).contiguous()
        B, Lf, _ = feat.shape
        cbs = fpc_codebooks.from_cfg(cfg, dev) if qtz else None
        if qtz and (cbs.arrays["vq"] is None or cbs.arrays["scl"] is None):
            raise FileNotFoundError("cfg['cb_path'] and cfg['scl_cb_path'] are required when qtz is set")
        weights = self.packed_weights(dev)
        shapes = {"c_in": ((B, Lf, 20), torch.float32), "r": ((B, Lf, 18), torch.float32),
                  "r_qtz": ((B, Lf, 18), torch.float32), "r_under": ((B, Lf, 18), torch.float32),
                  "ind1": ((B, Lf, 1), torch.float32), "ind2": ((B, Lf, 1), torch.float32),
                  "idx": ((B, Lf, 4), torch.int32)}
        res = dict(out) if out else {}
        for k, (shp, dt) in shapes.items():
            if k == "r_under" and not want_under and k not in res:
                continue
            t = res.get(k)
            if t is None:
                t = torch.empty(shp, dtype=dt).pin_memory()
            if t.is_cuda or t.dtype != dt or tuple(t.shape) != shp or not t.is_contiguous():
                raise ValueError("out[%r] must be a contiguous CPU %s tensor of shape %r" % (k, dt, shp))
            res[k] = t
        older = getattr(self, "_host_keep_older", None)
        if older is not None:
            older[4].synchronize()           # the call before the previous one has finished: its buffers may go
        self._host_keep_older = getattr(self, "_host_keep", None)
        with torch.cuda.device(dev):
            need = N.lib().fpc_encode_host_workspace_bytes(B, Lf, self.precision)
            ws = getattr(self, "_host_ws", None)
            if ws is None or ws.device != dev or ws.numel() < need:
                ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
                self._host_ws = ws
            io = N.EncodeHostIO()
            io.h_feat = feat.data_ptr()
            io.B, io.L = B, Lf
            io.l1, io.l2 = float(l1), float(l2)
            io.qtz = 1 if qtz else 0
            for k in shapes:
                setattr(io, "h_" + k, res[k].data_ptr() if k in res else None)
            if want_hist:
                h = res.get("hist")
                if h is None or h.dtype != torch.int64 or h.numel() != N.HIST_TOTAL or h.is_cuda:
                    h = torch.zeros(N.HIST_TOTAL, dtype=torch.int64).pin_memory()
                res["hist"] = h
                res["codebooks"] = cbs
                io.h_hist = h.data_ptr()
            N.check(N.lib().fpc_encode_host(weights.data_ptr(), cbs.ptr() if cbs is not None else None,
                                            ctypes.byref(io), self.precision, int(chunks), ws.data_ptr(), ws.numel(),
                                            N.current_stream(dev)), "fpc_encode_host")
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(dev))
        # The copies read / write the host tensors after this call returns.  They are kept alive here for two more calls
        # (the second of which waits on `done` before dropping them -- by then it has long completed, so the host
        # still enqueues one call ahead of the GPU), so a caller that discards a result without synchronising cannot
        # free pinned memory under a running DMA.
        self._host_keep = (cbs, feat, weights, res, done)
        return res

    @staticmethod
    def host_cb_tot(res):
        """cb_tot (wavernn.py:189,221-240) of an encode_host() result; call after synchronising the stream."""
        cbs, h = res.get("codebooks"), res["hist"].numpy()
        if cbs is None:
            return [0, 0, 0, 0, 0]
        out = []
        for off, n in zip(N.HIST_OFFSETS, cbs.hist_sizes()):
            t = h[off:off + n].astype(np.float64)
            out.append(t if (n > 0 and t.sum() > 0) else 0)
        return out

    def histograms(self, res):
        """cb_tot (wavernn.py:189,221-240): five usage tables from the index record."""
        cbs = res.codebooks[0]
        dev = res.idx.device
        hist = torch.empty(N.HIST_TOTAL, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib().fpc_index_histogram(res.idx.data_ptr(), res.idx.shape[0] * res.idx.shape[1],
                                                hist.data_ptr(), N.current_stream(dev)), "fpc_index_histogram")
        h = hist.cpu().numpy()
        sizes = cbs.hist_sizes()
        out = []
        for off, n in zip(N.HIST_OFFSETS, sizes):
            t = h[off:off + n].astype(np.float64)
            out.append(t if (n > 0 and t.sum() > 0) else 0)   # never-hit tables stay the int 0 of :189
        return out

    def encoder(self, cfg, feat, mask, l1, l2, vq_quantize=None, scl_quantize=None, qtz=True):
        """Closed-loop encode.  `vq_quantize` / `scl_quantize` are accepted for call-site
        compatibility (synthesis_qtz.py:151, generate_qtz_features.py:57); the quantisers of
        quantization/vq_func.py are fused into the kernel, so the callables are not invoked."""
        res = self.encode_device(cfg, feat, mask, l1, l2, qtz=bool(qtz))
        cb_tot = self.histograms(res) if qtz else [0, 0, 0, 0, 0]
        return res.c_in, res.r, res.r_qtz, res.r_under, res.ind1, res.ind2, cb_tot

    # ------------------------------------------------------------------ receiver side (wavernn.py:367-379)
    def decoder(self, cfg, feat, r):
        """Replays the recurrence from a quantised residual: c[t] = predictor(c[t-1]) + r[t],
        pitch dims copied from `feat`.  (The reference's decoder is not runnable -- h1/h2
        undefined, wavernn.py:375 -- this follows the encoder's feedback equation :242, so
        decoder(cfg, feat, r_qtz) reproduces encoder()'s c_in bit for bit.)"""
        N.require_cuda()
        if not (isinstance(feat, torch.Tensor) and feat.is_cuda):
            raise N.FpcError("Wavernn.decoder needs CUDA tensors: no CPU fallback")
        dev = feat.device
        r = torch.as_tensor(r).detach().to(device=dev, dtype=torch.float32).contiguous()
        B, Lf, _ = r.shape
        pitch = feat.detach()[:, :, -2:].to(torch.float32).contiguous()
        out = torch.empty((B, Lf, 20), dtype=torch.float32, device=dev)
        weights = self.packed_weights(dev)
        with torch.cuda.device(dev):
            N.check(N.lib().fpc_decode(weights.data_ptr(), r.data_ptr(), pitch.data_ptr(), B, Lf, out.data_ptr(),
                                       self.precision, None, 0, N.current_stream(dev)), "fpc_decode")
            torch.cuda.current_stream(dev).synchronize()
        return out
