"""Multi-GPU parity check, launched as one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

* closed-loop encode: every rank encodes its utterance shard; the gathered result must be
  BIT-IDENTICAL to rank 0 encoding all utterances alone (utterances are independent and seeded
  per utterance id; there is no collective on the data path), and equal to the oracle on a sample.
* k-means: every rank holds a shard of the residual vectors; after the NCCL all-reduce of the
  per-centroid sums/counts every rank must hold the same codebook, equal (1e-12 rel) to the
  single-GPU result and to the oracle; assignment indices must be identical.
(Not collected by pytest: it needs torchrun.  The gloo / world_size-2 CPU twin of the reduce step
is tests/test_host.py::test_kmeans_reduce_world_size_2_gloo.)
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "feature-predictor-for-speech-codec_b200"), os.path.join(ROOT, "oracle")]

import fpc_dist  # noqa: E402
import fpc_synth as S  # noqa: E402
import oracle as O  # noqa: E402
from models.wavernn import Wavernn  # noqa: E402
from quantization import cb_func  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    # ---------------- encode ----------------
    U, L, l1, l2 = 301, 50, 0.25, 2.1
    sd = S.make_state_dict(0)
    model = Wavernn(20, 384, 128, 18).eval()
    model.load_state_dict(sd)
    model = model.to(dev)
    cbs = S.make_codebooks(0)
    with tempfile.TemporaryDirectory(prefix="fpc_mg_%d_" % rank) as d:
        cfg = S.save_codebooks(cbs, d)
        first, cnt = fpc_dist.shard_range(U, rank, world)
        feat = S.make_features(cnt, L, first_utt=first)
        with torch.no_grad():
            out = model.encoder(cfg, torch.from_numpy(feat).to(dev), None, l1, l2, None, None, True)
        idx = model.last_result.idx
        sizes = [fpc_dist.shard_range(U, r, world)[1] for r in range(world)]
        gathered = {}
        for name, t in (("c_in", out[0]), ("r_qtz", out[2]), ("idx", idx)):
            parts = [torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for n in sizes]
            dist.all_gather(parts, t.contiguous())
            gathered[name] = torch.cat(parts, 0)
        hist = fpc_dist.merge_histograms(out[6])
        if rank == 0:
            full = S.make_features(U, L, first_utt=0)
            with torch.no_grad():
                ref = model.encoder(cfg, torch.from_numpy(full).to(dev), None, l1, l2, None, None, True)
            assert torch.equal(gathered["c_in"], ref[0]), "sharded c_in differs from single-GPU"
            assert torch.equal(gathered["r_qtz"], ref[2])
            assert torch.equal(gathered["idx"], model.last_result.idx)
            for a, b in zip(hist, ref[6]):
                assert np.array_equal(np.asarray(a), np.asarray(b)), "merged histograms differ"
            sample = [0, 150, 151, 300]
            ora = O.encode(O.weights_from_state_dict(sd),
                           O.Codebooks(cbs["cb_path"], cbs["scl_cb_path"], cbs["bl_cb_path"], cbs["bl_scl_cb_path"]),
                           full[sample], l1, l2)
            assert np.array_equal(gathered["idx"][sample].cpu().numpy(), ora["idx"])
            print("encode: %d utterances over %d ranks bit-identical to 1 GPU and to the oracle sample" % (U, world))

    # ---------------- k-means ----------------
    N, K = 200_003, 64
    data = S.make_kmeans_data(N, seed=41, n_components=48)
    cb0 = np.random.Generator(np.random.Philox(key=42)).standard_normal((K, 17)) * 0.1
    first, cnt = fpc_dist.shard_range(N, rank, world)
    shard = torch.from_numpy(data[first:first + cnt]).to(dev)
    cb = torch.from_numpy(cb0).to(dev)
    for _ in range(3):
        cb, stats, n_total = cb_func.update_device(shard, cb)
    assert n_total == N
    mine = cb.cpu().numpy()
    allcb = [torch.empty_like(cb) for _ in range(world)]
    dist.all_gather(allcb, cb)
    for r in range(world):
        assert torch.equal(allcb[r], cb), "ranks disagree on the codebook after the all-reduce"
    if rank == 0:
        ref = cb0
        for _ in range(3):
            ref = O.kmeans_update(data, ref)
        np.testing.assert_allclose(mine, ref, rtol=1e-11, atol=1e-300)
        print("k-means: 3 sharded Lloyd iterations on %d ranks match the oracle to 1e-11; ranks agree bit for bit" % world)
    # ordered sums (data order inside every shard, then the all-reduce): the same bits on every rank and on every run
    runs = []
    for _ in range(2):
        co = torch.from_numpy(cb0).to(dev)
        for _ in range(3):
            co, _, _ = cb_func.update_device(shard, co, ordered=True)
        runs.append(co)
    assert torch.equal(runs[0], runs[1]), "ordered update: two runs differ"
    allco = [torch.empty_like(co) for _ in range(world)]
    dist.all_gather(allco, runs[0])
    assert all(torch.equal(a, runs[0]) for a in allco), "ordered update: ranks disagree"
    if rank == 0:
        np.testing.assert_allclose(runs[0].cpu().numpy(), ref, rtol=1e-11, atol=1e-300)
        print("k-means ordered: reproducible run to run, ranks agree, oracle to 1e-11")
    # seeded vq_train: jitter broadcast from rank 0
    np.random.seed(1234 + rank)          # ranks deliberately seeded differently: rank 0's draw must win
    small = torch.from_numpy(data[first:first + cnt][:2000]).to(dev)
    trained = cb_func.vq_train(small, np.zeros((6, 17)), 6)
    t = torch.from_numpy(trained).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    assert all(torch.equal(p, t) for p in parts), "vq_train: ranks ended with different codebooks"
    if rank == 0:
        print("vq_train: ranks agree; done")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
