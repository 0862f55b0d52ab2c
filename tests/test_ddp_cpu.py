"""World-size-2 gloo test of the data-parallel host logic (flat gradient buffer, single
all-reduce, global-norm clip, AdamW contract).  The CUDA AdamW kernel is replaced by its
plain-torch statement here; the kernel itself is checked against torch.optim.AdamW in
tests/test_kernels_gpu.py."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def _tiny_model():
    from transformers import Blip2Config
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from eilev_b200.train import freeze_for_recipe
    fx = torch.load(GOLDEN / "tiny_opt.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    m = VideoBlipForConditionalGeneration(cfg)
    m.load_state_dict(fx["state_dict"])
    freeze_for_recipe(m)
    return m


def _fake_grads(trainer, seed):
    g = torch.Generator().manual_seed(seed)
    for _, p in trainer.flat.named:
        p.grad.copy_(torch.randn(p.shape, generator=g) * 3.0)  # large: the clip must engage


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eilev_b200.train import DataParallelTrainer, torch_adamw_reference
    m = _tiny_model()
    tr = DataParallelTrainer(m, lr=1e-2, weight_decay=0.05, max_grad_norm=1.0, grad_accum=2,
                             update_fn=torch_adamw_reference)
    assert tr.world == world
    for step in range(2):
        _fake_grads(tr, 100 * step + rank)
        tr.optimizer_step()
    torch.save({"params": tr.flat.params.clone(), "norm": tr.last_grad_norm.clone(),
                "named": {n: p.detach().clone() for n, p in tr.flat.named}}, Path(out_dir) / f"rank{rank}.pt")
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_mean_gradient(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["params"], r1["params"])  # replicas stay bit-identical

    # single-process statement: mean of the two ranks' gradients, clip_grad_norm_, AdamW
    from eilev_b200.train import DataParallelTrainer, torch_adamw_reference
    m = _tiny_model()
    tr = DataParallelTrainer(m, lr=1e-2, weight_decay=0.05, max_grad_norm=1.0, grad_accum=2,
                             update_fn=torch_adamw_reference)
    params = [p for _, p in tr.flat.named]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in params]
    from eilev_b200.train import no_weight_decay
    decay = [q for (n, _), q in zip(tr.flat.named, ref) if not no_weight_decay(n)]
    plain = [q for (n, _), q in zip(tr.flat.named, ref) if no_weight_decay(n)]
    assert decay and plain
    # HF Trainer's two parameter groups: LayerNorm parameters and biases are not decayed
    opt = torch.optim.AdamW([dict(params=decay, weight_decay=0.05), dict(params=plain, weight_decay=0.0)], lr=1e-2)
    for step in range(2):
        grads = []
        for rank in range(world):
            _fake_grads(tr, 100 * step + rank)
            grads.append([p.grad.clone() for p in params])
        for q, g0, g1 in zip(ref, *grads):
            q.grad = (g0 + g1) / world
        norm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
    assert abs(float(norm) - float(r0["norm"])) < 1e-3 * float(norm)
    for (n, _), q in zip(tr.flat.named, ref):
        assert torch.allclose(r0["named"][n], q.detach(), atol=1e-6, rtol=1e-5), n


def test_flat_buffers_alias_params_and_grads():
    from eilev_b200.train import FlatBuffers
    from eilev_b200.engine.qformer import qformer_param_list
    m = _tiny_model()
    fb = FlatBuffers(qformer_param_list(m))
    assert len(fb.named) == 47
    for (n, p), off in zip(fb.named, fb.offsets):
        assert p.data_ptr() == fb.params.data_ptr() + 4 * off, n
        assert p.grad.data_ptr() == fb.grads.data_ptr() + 4 * off, n
        assert off % 4 == 0
    fb.grads.fill_(2.0)
    assert all(float(p.grad.min()) == 2.0 for _, p in fb.named)
    fb.zero_grad()
    assert float(fb.grads.abs().sum()) == 0.0


def test_weight_decay_groups_match_hf_trainer():
    """The decayed / undecayed split of the flat buffer equals Trainer.get_decay_parameter_names
    (what train_v2.py's Trainer builds its AdamW groups from) on the trainable tensors."""
    from transformers.trainer_pt_utils import get_parameter_names
    from eilev_b200.train import FlatBuffers, no_weight_decay
    from eilev_b200.engine.qformer import qformer_param_list
    from transformers import Blip2Config, Blip2QFormerModel
    fx = torch.load(GOLDEN / "tiny_opt.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    hf_qformer = Blip2QFormerModel(cfg.qformer_config)  # the module the reference trains (v2.py:116)
    hf_decay = {"qformer." + n for n in get_parameter_names(hf_qformer, [torch.nn.LayerNorm]) if "bias" not in n}
    hf_all = {"qformer." + n for n, _ in hf_qformer.named_parameters()}
    m = _tiny_model()
    fb = FlatBuffers(qformer_param_list(m))
    names = [n for n, _ in fb.named]
    ours_decay = {n for n in names if not no_weight_decay(n)}
    for n in names:
        if n in hf_all:
            assert (n in hf_decay) == (n in ours_decay), n
    assert "query_tokens" in ours_decay and "language_projection.weight" in ours_decay
    assert "language_projection.bias" not in ours_decay
    # decayed tensors are one contiguous prefix of the flat buffer
    cut = fb.decay_numel
    for (n, p), off in zip(fb.named, fb.offsets):
        assert (off < cut) == (not no_weight_decay(n)), n
    assert 0 < cut < fb.numel


def test_stale_graph_pool_flag_forces_the_repack_graph():
    """after an optimizer step every same-shape micro-step must replay the re-pack graph
    first, even when eager (other-shape) micro-steps came in between."""
    from eilev_b200.train import DataParallelTrainer, torch_adamw_reference

    class FakeGraph:
        def __init__(self, log, name):
            self.log, self.name = log, name

        def replay(self):
            self.log.append(self.name)

    m = _tiny_model()
    tr = DataParallelTrainer(m, lr=1e-2, grad_accum=2, update_fn=torch_adamw_reference)
    log = []
    tr._static = {"x": torch.zeros(2, 3)}
    tr._graph, tr._graph_warm = FakeGraph(log, "repack"), FakeGraph(log, "warm")
    tr._static_loss = tr._static_loss_warm = torch.zeros(())
    tr._fwd_bwd = lambda batch: (log.append("eager"), torch.zeros(()))[1]
    same, other = {"x": torch.ones(2, 3)}, {"x": torch.ones(2, 5)}
    tr.micro_step(same)    # first after capture: re-pack
    tr.micro_step(same)    # warm; optimizer step follows (accum 2)
    tr.micro_step(other)   # eager: the graphs' pool is still stale
    tr.micro_step(same)    # must re-pack, not replay the warm graph; optimizer step follows
    tr.micro_step(same)    # re-pack again
    tr.micro_step(same)    # warm
    assert log == ["repack", "warm", "eager", "repack", "repack", "warm"], log


def test_linear_schedule_matches_hf():
    from transformers import get_linear_schedule_with_warmup
    from eilev_b200.train import linear_schedule
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1e-5)
    sch = get_linear_schedule_with_warmup(opt, 3, 20)
    f = linear_schedule(1e-5, 20, 3)
    for step in range(1, 21):
        assert abs(f(step) - opt.param_groups[0]["lr"]) < 1e-12, step
        opt.step()
        sch.step()


def _gather_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eilev_b200.train import gather_generated
    # rank 0: 3 rows of 4 new tokens; rank 1: the short last batch, 1 row of 6 tokens
    ids = torch.arange(12).view(3, 4) + 10 if rank == 0 else torch.arange(6).view(1, 6) + 100
    torch.save(gather_generated(ids, pad_token_id=1), Path(out_dir) / f"gathered{rank}.pt")
    dist.destroy_process_group()


def test_gather_generated_pads_and_concatenates_in_rank_order(tmp_path):
    """generate_narration_texts.py:120-124 (accelerate pad_across_processes + gather) on 2 gloo ranks."""
    from eilev_b200.train import gather_generated
    world = 2
    port = 31500 + os.getpid() % 2000
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "gathered0.pt"), torch.load(tmp_path / "gathered1.pt")
    want = torch.full((4, 6), 1, dtype=torch.long)
    want[:3, :4] = torch.arange(12).view(3, 4) + 10
    want[3] = torch.arange(6) + 100
    assert torch.equal(g0, want) and torch.equal(g1, want)
    solo = torch.arange(6).view(2, 3)
    assert gather_generated(solo, 0) is solo  # no process group: identity
