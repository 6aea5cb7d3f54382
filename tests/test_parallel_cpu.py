"""Host-side logic of the multi-GPU path on CPU: bucket planning, exact global-mean loss from
all-reduced sums, request sharding + gather — exercised with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ergm_b200 import parallel
from oracle import ergm_oracle as O


def _entries(cfg):
    off, ent = 0, {}
    for name, shape in O.param_shapes(cfg):
        n = 1
        for s in shape:
            n *= s
        ent[name] = (off, n, shape)
        off += (n + 63) // 64 * 64
    return ent, off


def test_plan_buckets_cover_flat_buffer_in_backward_order():
    cfg = O.OracleConfig(vocab_size=1024, n_positions=64, n_embd=128, n_layer=6, n_head=2)
    ent, total = _entries(cfg)
    per_layer = ent["transformer.h.1.ln_1.weight"][0] - ent["transformer.h.0.ln_1.weight"][0]
    buckets = parallel.plan_buckets(ent, cfg.n_layer, bucket_bytes=2 * per_layer * 4)
    # contiguous, non-overlapping, covering [0, total), descending addresses, embeddings last
    assert buckets[-1][0] == -1 and buckets[-1][1] == 0
    assert buckets[0][2] == total
    for (t0, lo0, hi0), (t1, lo1, hi1) in zip(buckets[:-1], buckets[1:]):
        assert lo0 == hi1 and lo0 < hi0
    triggers = [t for t, _, _ in buckets[:-1]]
    assert triggers == sorted(triggers, reverse=True) and triggers[-1] == 0
    # every bucket but the last layer group is at least the requested size
    assert all((hi - lo) * 4 >= 2 * per_layer * 4 for _, lo, hi in buckets[:-2])


def test_shard_range_partitions():
    for n in (1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    # per-rank token losses with a different number of valid (non -100) labels per rank
    n_valid = 5 + 7 * rank
    tok_losses = torch.rand(n_valid) * 3
    emo_losses = torch.rand(4)
    sums = torch.tensor([tok_losses.sum(), float(n_valid), emo_losses.sum(), 4.0])
    dist.all_reduce(sums)
    g = parallel.global_loss_from_sums(sums.tolist())
    # gather all per-token losses to form the reference "concatenated batch" mean on rank 0
    all_tok = [torch.zeros(5 + 7 * r) for r in range(world)]
    all_emo = [torch.zeros(4) for _ in range(world)]
    for r in range(world):
        t = tok_losses.clone() if r == rank else torch.zeros(5 + 7 * r)
        dist.broadcast(t, r)
        all_tok[r] = t
        e = emo_losses.clone() if r == rank else torch.zeros(4)
        dist.broadcast(e, r)
        all_emo[r] = e
    want = torch.cat(all_tok).mean() + torch.cat(all_emo).mean()
    naive = tok_losses.mean() + emo_losses.mean()  # what a per-rank mean would give
    # sharded generation gather
    n = 7
    lo, hi = parallel.shard_range(n, rank, world)
    local = torch.arange(lo, hi).view(-1, 1).repeat(1, 3)
    full = parallel.gather_rows(local, n, rank, world)
    ok_gather = torch.equal(full, torch.arange(n).view(-1, 1).repeat(1, 3))
    # bucketed all-reduce of a fake flat gradient equals the single all-reduce
    cfg = O.OracleConfig(vocab_size=256, n_positions=32, n_embd=128, n_layer=3, n_head=2)
    ent, total = _entries(cfg)
    flat = torch.arange(total, dtype=torch.float32) * (rank + 1)
    ref = flat.clone()
    dist.all_reduce(ref)
    works = []
    for _, lo_, hi_ in parallel.plan_buckets(ent, 3, 1 << 16):
        works.append(dist.all_reduce(flat[lo_:hi_], async_op=True))
    for w in works:
        w.wait()
    q.put((rank, abs(g - want.item()) < 1e-6, abs(naive.item() - want.item()) > 1e-4, ok_gather, torch.equal(flat, ref)))
    dist.destroy_process_group()


def test_world2_gloo_global_mean_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, exact, naive_differs, ok_gather, ok_bucket in res:
        assert exact, "global mean from all-reduced sums must equal the concatenated-batch mean"
        assert ok_gather and ok_bucket
    assert any(r[2] for r in res)


def test_last_bucket_is_the_embedding_bucket():
    """DataParallel defers the LAST bucket (wte + wpe, final only after the embedding backward) and runs
    AdamW on early_ranges() meanwhile: the plan must put exactly the embedding range last."""
    from ergm_b200.parallel import plan_buckets
    entries, off = {}, 0

    def add(name, n):
        nonlocal off
        entries[name] = (off, n, (n,))
        off += (n + 63) // 64 * 64

    add("transformer.wte.weight", 1000 * 64)
    add("transformer.wpe.weight", 100 * 64)
    for l in range(4):
        add("transformer.h.%d.ln_1.weight" % l, 64)
        add("transformer.h.%d.mlp.c_fc.weight" % l, 64 * 256)
    add("transformer.ln_f.weight", 64)
    add("emotion_head.weight", 7 * 64)
    for mb in (0.001, 0.05, 128):
        b = plan_buckets(entries, 4, int(mb * (1 << 20)))
        trig, lo, hi = b[-1]
        assert trig == -1 and lo == 0 and hi == entries["transformer.h.0.ln_1.weight"][0]
        covered = sorted((lo, hi) for _, lo, hi in b)
        assert covered[0][0] == 0 and covered[-1][1] == off
        assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))


def test_projection_parameters_ride_in_the_embedding_stage_buckets():
    """A3 extension (config.ergm_visual_dim): visual_proj / audio_proj are registered after emotion_head, but
    their gradients are written at the very END of the backward (after embed_bwd).  They must be in a
    trigger -1 bucket, never in the bucket the last layers trigger (which is reduced long before)."""
    cfg = O.OracleConfig(vocab_size=1024, n_positions=64, n_embd=128, n_layer=4, n_head=2, visual_dim=96, audio_dim=96)
    ent, total = _entries(cfg)
    for mb in (0.001, 0.05, 128):
        b = parallel.plan_buckets(ent, cfg.n_layer, int(mb * (1 << 20)))
        late = [(lo, hi) for t, lo, hi in b if t == -1]
        proj_lo = ent["visual_proj.weight"][0]
        assert (0, ent["transformer.h.0.ln_1.weight"][0]) in late and (proj_lo, total) in late
        for t, lo, hi in b:
            if t >= 0:  # layer-triggered buckets stop before the projections
                assert hi <= proj_lo
        assert ent["emotion_head.weight"][0] < proj_lo  # the head still rides with the last layers
        covered = sorted((lo, hi) for _, lo, hi in b)
        assert covered[0][0] == 0 and covered[-1][1] == total
        assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
