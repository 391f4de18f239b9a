"""Host-side logic that needs no GPU: state_dict compatibility with the reference's modules, API surface,
CPU-tensor rejection, word sharding and the final gather (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import paule_oracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_keys_and_seeded_init_match_reference(golden):
    import paule_b200 as P
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=720)
    # identical parameter creation order + naming => the reference's seeded weights, bit for bit
    assert [O.state_dict_digest(m) for m in (pred, emb, inv)] == list(golden["digest32"])
    opred, oemb, oinv = O.build_reference_models(0, 720)
    for mine, ref in ((pred, opred), (emb, oemb), (inv, oinv)):
        assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
        mine.load_state_dict(ref.state_dict())           # reference checkpoints load unchanged
    assert P.InverseModel is P.InverseModelMelTimeSmoothResidual and P.PAULE is P.Paule


def test_default_constructor_arguments_match_reference():
    import inspect
    import paule_b200 as P
    sig = inspect.signature(P.ForwardModel.__init__).parameters
    assert [sig[k].default for k in ("input_size", "output_size", "hidden_size", "num_lstm_layers", "apply_half_sequence")] \
        == [30, 60, 180, 4, True]                                  # paule/models.py:335-339
    sig = inspect.signature(P.EmbeddingModel.__init__).parameters
    assert [sig[k].default for k in ("input_size", "output_size", "hidden_size", "num_lstm_layers",
                                     "post_upsampling_size", "dropout")] == [60, 300, 720, 1, 0, 0]   # :421-427
    sig = inspect.signature(P.Paule.plan_resynth).parameters
    for k, v in dict(learning_rate_planning=0.01, initialize_from="acoustic", objective="acoustic", n_outer=5,
                     n_inner=24, continue_learning=True, log_ii=1, log_semantics=True, log_cps=False).items():
        assert sig[k].default == v and sig[k].kind == inspect.Parameter.KEYWORD_ONLY     # paule/paule.py:391-414


def test_planning_results_fields():
    import paule_b200 as P
    f = P.PlanningResults._fields
    assert len(f) == 33                                               # paule/paule.py:57
    assert f[:2] == ("planned_cp", "initial_cp") and f[18:23] == ("planned_loss_steps", "planned_mel_loss_steps",
                                                                    "vel_loss_steps", "jerk_loss_steps",
                                                                    "pred_semvec_loss_steps")


def test_cpu_inputs_fail_loudly():
    import paule_b200 as P
    from paule_b200 import _lib
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=16)
    with pytest.raises(_lib.PauleB200Error):
        pred(torch.zeros(1, 20, 30))
    with pytest.raises(_lib.PauleB200Error):
        P.Paule(pred_model=pred, device=torch.device("cpu"))


def test_shard_bounds_cover_the_word_axis():
    from paule_b200.distributed import shard_bounds
    for n in (0, 1, 7, 64, 2048, 2049):
        for ws in (1, 2, 4, 8):
            spans = [shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


class _FakePlanner:
    """Stands in for BatchPlanner on CPU: a deterministic, word-independent 'plan'."""

    def __init__(self, cp, mel, sv):
        self.cp = cp.clone()
        self.mel = mel
        self.log = []

    def step(self, n):
        for _ in range(n):
            self.log.append(self.cp.flatten(1).pow(2).mean(1) + self.mel.flatten(1).mean(1))
            self.cp = 0.9 * self.cp

    def planned_cp(self):
        return self.cp

    def losses(self):
        return {"total": torch.stack(self.log)}


def _worker(rank, ws, port, n_words, out_q):
    sys.path.insert(0, REPO)
    from paule_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    g = torch.Generator().manual_seed(0)
    cp = torch.rand(n_words, 20, 30, generator=g)
    mel = torch.rand(n_words, 10, 60, generator=g)
    cps, loss = D.plan_sharded(_FakePlanner, cp, mel, None, 3, gather_dst=None)
    cps0, loss0 = D.plan_sharded(_FakePlanner, cp, mel, None, 3, gather_dst=0)
    assert (cps0 is None) == (rank != 0)
    out_q.put((rank, cps.numpy(), loss.numpy(), None if cps0 is None else cps0.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_words", [6, 7])
def test_sharded_plan_equals_single_rank_gloo_world2(n_words):
    """N-rank sharded result == 1-rank result (words are independent); ragged shard when n_words is odd."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_words
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_words, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    cp = torch.rand(n_words, 20, 30, generator=g)
    mel = torch.rand(n_words, 10, 60, generator=g)
    ref = _FakePlanner(cp, mel, None)
    ref.step(3)
    for rank, cps, loss, cps0 in res:
        np.testing.assert_array_equal(cps, ref.planned_cp().numpy())
        np.testing.assert_array_equal(loss, ref.losses()["total"].numpy())
        if rank == 0:
            np.testing.assert_array_equal(cps0, ref.planned_cp().numpy())


class _FakePaule:
    """Stands in for paule_b200.Paule in the CPU test of plan_resynth_sharded: plan_resynth on host arrays -> results object,
    and a `last_planner` with planned_cp() / losses() (device tensors on the GPU box, CPU tensors here)."""

    def plan_resynth(self, *, target_acoustic, initial_cp, target_semvec=None, n_inner=3, **kw):
        self.last_planner = _FakePlanner(torch.as_tensor(initial_cp), torch.as_tensor(target_acoustic), None)
        self.last_planner.step(n_inner)
        return {"planned_cp": self.last_planner.planned_cp().numpy(), "n_words": len(target_acoustic)}


def _resynth_worker(rank, ws, port, n_words, out_q):
    sys.path.insert(0, REPO)
    from paule_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    g = torch.Generator().manual_seed(0)
    cp = torch.rand(n_words, 20, 30, generator=g).numpy()
    mel = torch.rand(n_words, 10, 60, generator=g).numpy()
    r = D.plan_resynth_sharded(_FakePaule(), target_acoustic=mel, initial_cp=cp, n_inner=3)
    r0 = D.plan_resynth_sharded(_FakePaule(), target_acoustic=mel, initial_cp=cp, n_inner=3, gather_dst=0)
    assert (r0.planned_cp is None) == (rank != 0)
    out_q.put((rank, r.planned_cp, r.planned_loss_steps, r.word_range, r.local["n_words"]))
    dist.barrier()
    dist.destroy_process_group()


def test_plan_resynth_sharded_gathers_the_whole_job_gloo_world2():
    """distributed.plan_resynth_sharded (what bench.py times at N > 1): every rank plans its shard through the Paule API, one
    final all_gather returns the whole job -- equal to the job planned on one rank."""
    n_words = 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_resynth_worker, args=(r, 2, port, n_words, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    cp = torch.rand(n_words, 20, 30, generator=g)
    mel = torch.rand(n_words, 10, 60, generator=g)
    ref = _FakePlanner(cp, mel, None)
    ref.step(3)
    for rank, cps, loss, rng, n_local in res:
        np.testing.assert_array_equal(cps, ref.planned_cp().numpy())
        np.testing.assert_array_equal(loss, ref.losses()["total"].numpy())
        assert rng == ((0, 4) if rank == 0 else (4, 7)) and n_local == rng[1] - rng[0]


# ---- ragged jobs: length-bucketed sharding --------------------------------------------------------------------------
def test_length_buckets_balance_cost_and_cover_every_word():
    from paule_b200.distributed import length_buckets
    rng = np.random.RandomState(0)
    for ws in (1, 2, 3, 8):
        for n in (0, 1, 5, 64, 257):
            lengths = (2 * rng.randint(7, 600, size=n)).tolist()
            b = length_buckets(lengths, ws)
            assert len(b) == ws and sorted(i for r in b for i in r) == list(range(n))
            if n == 0:
                continue
            cost = [max(lengths[i] for i in r) * len(r) if r else 0 for r in b]
            # every rank holds a contiguous run of the length-sorted words ...
            flat = [lengths[i] for r in b for i in r]
            assert flat == sorted(lengths, reverse=True)
            # ... and no rank costs more than the even share plus one longest word
            assert max(cost) <= sum(l for l in lengths) / ws + 2 * max(lengths) * max(1, n // (4 * ws)) or ws == 1
    assert length_buckets([100, 20, 20, 20, 20, 20], 2) == [[0], [1, 2, 3, 4, 5]]


class _FakeRaggedPlanner(_FakePlanner):
    def __init__(self, cp, mel, lengths):
        super().__init__(cp, mel, None)
        self.lengths = lengths
        m = torch.zeros(cp.shape[:2])
        for b, L in enumerate(lengths):
            m[b, :L] = 1
        self.mask = m

    def step(self, n):
        for _ in range(n):
            per = (self.cp.pow(2).mean(2) * self.mask).sum(1) / self.mask.sum(1)
            self.log.append(per)
            self.cp = 0.9 * self.cp


def _ragged_job(n_words):
    g = torch.Generator().manual_seed(1)
    lengths = [2 * int(v) for v in torch.randint(7, 40, (n_words,), generator=g)]
    cps = [torch.rand(L, 30, generator=g) for L in lengths]
    mels = [torch.rand(L // 2, 60, generator=g) for L in lengths]
    return lengths, cps, mels


def _ragged_worker(rank, ws, port, n_words, out_q):
    sys.path.insert(0, REPO)
    from paule_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    _, cps, mels = _ragged_job(n_words)
    planned, loss = D.plan_sharded_ragged(_FakeRaggedPlanner, cps, mels, 3)
    out_q.put((rank, [p.numpy() for p in planned], loss.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_sharded_plan_equals_words_planned_alone_gloo_world2():
    n_words = 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ragged_worker, args=(r, 2, port, n_words, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lengths, cps, mels = _ragged_job(n_words)
    for rank, planned, loss in res:
        for b in range(n_words):
            solo = _FakeRaggedPlanner(cps[b][None], mels[b][None], [lengths[b]])
            solo.step(3)
            np.testing.assert_array_equal(planned[b], solo.planned_cp()[0].numpy())
            np.testing.assert_allclose(loss[:, b], torch.stack(solo.log)[:, 0].numpy(), rtol=1e-6)


def test_mel_embedding_model_state_dict_matches_reference_keys():
    """MelEmbeddingModelMelSmoothResidualUpsampling (paule/models.py:362-409): same parameter names and seeded init as the
    reference (golden generated by tests/golden/make_mel_embedder_golden.py from the reference class)."""
    import hashlib
    import paule_b200 as P
    g = np.load(os.path.join(REPO, "tests", "golden", "mel_embedder_golden.npz"))
    torch.manual_seed(7)
    m = P.MelEmbeddingModelMelSmoothResidualUpsampling(hidden_size=96, num_lstm_layers=2, post_upsampling_size=256)
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(g["keys"])
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode()); h.update(sd[k].detach().cpu().float().numpy().tobytes())
    assert h.hexdigest() == str(g["digest"])


def test_generator_matches_the_reference_output():
    """paule_b200.models.Generator (prologue-only, library ops) == the reference's Generator on seeded weights
    (tests/golden/make_generator_golden.py); pure torch, so it runs on the CPU here."""
    import os
    import numpy as np
    import torch
    from paule_b200.models import Generator
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "generator_golden.npz")))
    for tag, osz, length in (("cp", 30, 46), ("mel", 60, 23)):
        torch.manual_seed(4)
        gen = Generator(output_size=osz).eval()
        with torch.no_grad():
            y = gen(torch.from_numpy(g[f"{tag}_noise"]), length, torch.from_numpy(g[f"{tag}_vec"]))
        assert y.shape == (2, length, osz)
        np.testing.assert_allclose(y.numpy(), g[f"{tag}_y"], atol=1e-6)
