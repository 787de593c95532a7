"""Drop-in for /root/reference/src/ceps2lpc/ceps2lpc_vct.py: `ceps2lpc_v(cepstrum) -> (e, lpc, rc)`.

The reference runs IDCT / band interpolation on the CPU and the Levinson recursion frame by frame in a
Python loop (:150-154); here one CUDA thread per frame does the whole chain (`fpc_ceps2lpc`,
csrc/fpc_ceps2lpc.cu).  Return values follow the reference: `lpc` (N,16) float32, and `e` / `rc` of the
LAST frame only (they are loop variables there, :151-153,162).  The call sites pass a CPU tensor
(`c_in.reshape(-1, C).cpu()`, synthesis_qtz.py:159); a CUDA tensor is accepted too and then `lpc` stays on
the device.  No CPU fallback.
"""
import torch

import fpc_native as N

NB_BANDS = 18
LPC_ORDER = 16


def ceps2lpc_device(cepstrum):
    """(N, C>=18) CUDA float32 -> (lpc (N,16), err (N,), rc (N,16)) on the device, asynchronous."""
    N.require_cuda()
    if not (isinstance(cepstrum, torch.Tensor) and cepstrum.is_cuda):
        raise N.FpcError("ceps2lpc_device needs a CUDA tensor")
    if cepstrum.dim() != 2 or cepstrum.shape[1] < NB_BANDS:
        raise ValueError("cepstrum must be (frames, >=18), got %r" % (tuple(cepstrum.shape),))
    x = cepstrum.detach().to(torch.float32).contiguous()
    n = x.shape[0]
    lpc = torch.empty((n, LPC_ORDER), dtype=torch.float32, device=x.device)
    err = torch.empty((n,), dtype=torch.float32, device=x.device)
    rc = torch.empty((n, LPC_ORDER), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().fpc_ceps2lpc(x.data_ptr(), n, x.shape[1], lpc.data_ptr(), err.data_ptr(), rc.data_ptr(),
                                     N.current_stream(x.device)), "fpc_ceps2lpc")
    return lpc, err, rc


def ceps2lpc_v(cepstrum):
    """ceps2lpc_vct.py:122-162."""
    N.require_cuda()
    t = torch.as_tensor(cepstrum)
    on_device = t.is_cuda
    lpc, err, rc = ceps2lpc_device(t if on_device else t.cuda())
    if t.shape[0] == 0:
        raise UnboundLocalError("ceps2lpc_v on zero frames: the reference's loop variables e/rc are unbound")
    e = err[-1]
    rc_last = rc[-1].to(torch.float64)
    if on_device:
        return e, lpc, rc_last
    return e.cpu(), lpc.cpu(), rc_last.cpu()
