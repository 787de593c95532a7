"""Codebook-usage entropy / bitrate report.

`cal_entropy` is the function of /root/reference/src/generate_qtz_features.py:94-101 (same in-place
normalisation and log2 with the 1e-20 guard); `bitrate_report` turns the five `cb_tot` tables returned by
`Wavernn.encoder` (wavernn.py:189,221-240) into bits per 10 ms frame and kbit/s, which is the figure the paper
quotes and the reference only prints table by table (:202).  Host-side NumPy over five small tables.
"""
import numpy as np


def cal_entropy(cb):
    """Entropy in bits of a usage histogram; normalises `cb` in place like the reference."""
    cb /= np.sum(cb)
    ent = np.sum(- cb * np.log2(cb + 1e-20))
    return ent


def bitrate_report(cb_tot, frame_ms=10.0):
    """cb_tot = [scl_above, scl_below, vq_stage1, vq_stage2, vq_below] (never-hit tables may be the int 0).

    Returns a dict with the entropy of every table, the fraction of frames coded by each branch, the expected
    index bits per frame (entropy-coded and fixed-length) and the two indicator bits, in bits/frame and kbit/s."""
    tabs = [None if np.isscalar(h) else np.asarray(h, dtype=np.float64) for h in cb_tot]
    cnt = [0.0 if t is None else float(t.sum()) for t in tabs]
    ent = [0.0 if (t is None or t.sum() == 0) else float(cal_entropy(t.copy())) for t in tabs]
    fixed = [0.0 if t is None else float(np.ceil(np.log2(max(len(t), 2)))) for t in tabs]
    n_c0 = cnt[0] + cnt[1]
    n_vq = cnt[2] + cnt[4]
    frames = max(n_c0, n_vq, 1.0)
    use = [cnt[0] / frames, cnt[1] / frames, cnt[2] / frames, cnt[3] / frames, cnt[4] / frames]
    bits_entropy = 2.0 + sum(u * e for u, e in zip(use, ent))
    bits_fixed = 2.0 + sum(u * f for u, f in zip(use, fixed))
    return {
        "frames": frames, "entropy_bits": ent, "fixed_bits": fixed, "usage": use,
        "bits_per_frame_entropy_coded": bits_entropy, "bits_per_frame_fixed_length": bits_fixed,
        "kbps_entropy_coded": bits_entropy / frame_ms, "kbps_fixed_length": bits_fixed / frame_ms,
    }
