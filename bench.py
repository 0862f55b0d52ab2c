"""Headline benchmark: clips/sec of the VideoBLIP fwd+bwd training step
(eilev-blip2-opt-2.7b, 16 in-context clips + 1 query clip x 8 frames, L = 976, bs 1 per GPU,
grad-accum 16, data parallel) on N B200s — BASELINE.json configs[1] / configs[2].

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # CPU arm: the oracle port of
                                                              # the reference on host cores

A "step" is one micro-step = forward + backward of one synthetic datapoint (17 clips); every
16th step also runs the gradient all-reduce + clip + fused AdamW.  Random-init weights of
the real architecture (sane seeded init) and synthetic frames: no network.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
os.environ.setdefault("HF_HUB_OFFLINE", "1")

import torch  # noqa: E402

CLIPS, FRAMES, QUERY_TOKENS, TEXT_PER_CLIP, TARGET_TOKENS = 17, 8, 32, 24, 12
GRAD_ACCUM = 16
# algorithmic FLOPs per datapoint: ViT fwd 70.82 + Q-Former 1.03/1.15 + proj
# + OPT fwd 5.32 / dgrad-only bwd 5.55
FLOPS_PER_DATAPOINT = 83.88e12


def full_config(dropout: float = 0.0, lm: str = "opt"):
    """eilev-blip2-opt-2.7b architecture (lm="opt") or eilev-blip2-flan-t5-xl (lm="t5",
    BASELINE configs[3]).  dropout = 0.1 is the checkpoint's / recipe's value (Q-Former hidden +
    attention-probs dropout, OPT hidden dropout; OPT attention_dropout 0)."""
    from transformers import Blip2Config

    if lm == "t5":
        text = dict(model_type="t5", d_model=2048, d_kv=64, d_ff=5120, num_layers=24, num_decoder_layers=24,
                    num_heads=32, vocab_size=32128, feed_forward_proj="gated-gelu", tie_word_embeddings=False,
                    decoder_start_token_id=0, pad_token_id=0, eos_token_id=1, dropout_rate=dropout,
                    relative_attention_num_buckets=32, relative_attention_max_distance=128)
        return Blip2Config(
            vision_config=dict(hidden_size=1408, intermediate_size=6144, num_hidden_layers=39,
                               num_attention_heads=16, patch_size=14, image_size=224, hidden_act="gelu",
                               layer_norm_eps=1e-6, qkv_bias=True),
            qformer_config=dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                                intermediate_size=3072, encoder_hidden_size=1408, cross_attention_frequency=2,
                                vocab_size=30522, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout),
            text_config=text, num_query_tokens=QUERY_TOKENS)
    return Blip2Config(
        vision_config=dict(hidden_size=1408, intermediate_size=6144, num_hidden_layers=39,
                           num_attention_heads=16, patch_size=14, image_size=224, hidden_act="gelu",
                           layer_norm_eps=1e-6, qkv_bias=True),
        qformer_config=dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                            intermediate_size=3072, encoder_hidden_size=1408, cross_attention_frequency=2,
                            vocab_size=30522, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout),
        text_config=dict(model_type="opt", hidden_size=2560, num_hidden_layers=32, ffn_dim=10240,
                         num_attention_heads=32, vocab_size=50272, max_position_embeddings=2048,
                         word_embed_proj_dim=2560, dropout=dropout, attention_dropout=0.0),
        num_query_tokens=QUERY_TOKENS)


def synthetic_batch(seed: int, clips: int = CLIPS, frames: int = FRAMES, pad_to: int = 8, lm: str = "opt"):
    """[bos] + clips x (32 pad-id slots + '\\n' + 24 text ids), right
    padded to a multiple of 8 (train_v2.py:214); labels = last 12 text tokens.
    lm="t5" (config 4, eilev/data/utils.py:200-217): no bos, pad id 0, '\\n' = 3, eos = 1 after
    the last prompt; labels = 12 target ids for the decoder."""
    g = torch.Generator().manual_seed(seed)
    px = torch.randn(clips, 3, frames, 224, 224, generator=g)
    if lm == "t5":
        ids, vm = [], []
        for _ in range(clips):
            ids += [0] * QUERY_TOKENS + [3] + torch.randint(4, 32000, (TEXT_PER_CLIP,), generator=g).tolist()
            vm += [1] * QUERY_TOKENS + [0] * (1 + TEXT_PER_CLIP)
        ids += [1]
        vm += [0]
        n = len(ids)
        pad = (-n) % pad_to
        return dict(
            pixel_values=px,
            input_ids=torch.tensor([ids + [0] * pad]),
            attention_mask=torch.tensor([[1] * n + [0] * pad]),
            video_input_mask=torch.tensor([vm + [0] * pad]),
            labels=torch.randint(4, 32000, (1, TARGET_TOKENS), generator=g),
        )
    ids, vm = [2], [0]
    for _ in range(clips):
        ids += [1] * QUERY_TOKENS + [50118] + torch.randint(4, 50000, (TEXT_PER_CLIP,), generator=g).tolist()
        vm += [1] * QUERY_TOKENS + [0] * (1 + TEXT_PER_CLIP)
    labels = [-100] * (len(ids) - TARGET_TOKENS) + ids[-TARGET_TOKENS:]
    n = len(ids)
    pad = (-n) % pad_to
    return dict(
        pixel_values=px,
        input_ids=torch.tensor([ids + [1] * pad]),
        attention_mask=torch.tensor([[1] * n + [0] * pad]),
        video_input_mask=torch.tensor([vm + [0] * pad]),
        labels=torch.tensor([labels + [-100] * pad]),
    )


def t5_flops_per_datapoint(l: int = 976, ld: int = TARGET_TOKENS) -> float:
    """ViT + Q-Former (as for OPT) + flan-t5-xl fwd and dgrad-only bwd (attention bwd 2.5x)."""
    d, inner, dff, v, layers = 2048, 2048, 5120, 32128, 24
    enc_lin = 2.0 * l * (3 * d * inner + inner * d + 3 * d * dff) * layers
    enc_att = 4.0 * l * l * inner * layers
    dec_lin = 2.0 * ld * (3 * d * inner + inner * d + 2 * d * inner + 3 * d * dff) * layers + 2.0 * ld * d * v
    ckv = 2.0 * l * d * 2 * inner * layers
    dec_att = (2.0 * ld * ld + 4.0 * ld * l) * inner * layers
    fwd = enc_lin + enc_att + dec_lin + ckv + dec_att
    bwd = enc_lin + dec_lin + ckv + 2.5 * (enc_att + dec_att)
    return 70.82e12 + 1.028e12 + 1.150e12 + 0.0064e12 * 2048 / 2560 + fwd + bwd


def workload_config(world: int, seq_len: int, cuda_graph, dropout: float = 0.1, lm: str = "opt"):
    name = ("eilev-blip2-flan-t5-xl 16-ctx x 8-frame fwd+bwd bs=1 per GPU (17 clips, encoder L=976, 12 target "
            "tokens)" if lm == "t5" else
            "eilev-blip2-opt-2.7b 16-ctx x 8-frame fwd+bwd bs=1 per GPU (17 clips, L=976)")
    cfg = {"workload": name + ", grad-accum 16 with all-reduce + AdamW every 16th step",
           "global_batch": world, "seq_len": seq_len, "parallelism": f"dp{world}",
           "weights": "random-init (seeded N(0,0.02))",
           "l2": "per-step working set (7.3 GB bf16 weights + activations) >> 126 MB L2",
           "dropout": (("recipe: p=%.2f in train mode (Q-Former hidden + attention probs; T5 dropout_rate at every "
                        "site of both stacks)" if lm == "t5" else
                        "recipe: p=%.2f in train mode (Q-Former hidden + attention probs, OPT hidden)") % dropout)
           if dropout > 0 else "off"}
    if cuda_graph is not None:
        cfg["cuda_graph"] = cuda_graph
    return cfg


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.rows, self.proc = index, [], None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def build_cpu_state_dict(cfg, seed: int = 1234):
    """Random-init fp32 state_dict with the reference's key layout, built without the GPU."""
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from oracle.videoblip_ref import sane_init_

    with torch.device("meta"):
        skeleton = VideoBlipForConditionalGeneration(cfg)
    sd = {}
    g = torch.Generator().manual_seed(seed)
    for k, v in skeleton.state_dict().items():
        if k == "language_model.lm_head.weight":
            continue
        lk = k.lower()
        if "layernorm" in lk or "layer_norm" in lk:
            sd[k] = torch.ones(v.shape) if k.endswith("weight") else torch.zeros(v.shape)
        else:
            sd[k] = torch.empty(v.shape).normal_(0.0, 0.02, generator=g)
    sd["language_model.lm_head.weight"] = sd["language_model.model.decoder.embed_tokens.weight"]
    del sane_init_
    return sd


def cpu_reference_step(sd, cfg, batch, trainable):
    """The reference path (oracle port of eilev/model/v2.py:132-252 + backward), fp32, on the
    host cores: fwd + bwd of `batch`, gradients for the trainable tensors only."""
    from oracle import videoblip_ref as R

    for k in trainable:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    out = R.videoblip_forward(sd, cfg, **batch)
    out["loss"].backward()
    return float(out["loss"])


def run_cpu_baseline(cfg, clips: int, steps: int, warmup: int, threads: int | None = None):
    """The oracle PORT on the host cores (kind = "port"): only used when baseline/_ref is absent."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = build_cpu_state_dict(cfg)
    trainable = [k for k in sd if k.startswith(("qformer.", "query_tokens", "language_projection."))]
    batch = synthetic_batch(1, clips=clips)
    for _ in range(warmup):
        cpu_reference_step(sd, cfg, batch, trainable)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, cfg, batch, trainable)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=clips / dt, unit="clips/s", cores=threads, kind="port",
                sample=f"{steps} x fwd+bwd of {clips} clip(s) x {FRAMES} frames (L={batch['input_ids'].shape[1]}), "
                       f"full-size model, fp32, oracle port of the reference",
                ms_per_step=dt * 1e3)


def run_reference(cfg, clips: int, steps: int, warmup: int, device: str = "cpu"):
    """The REAL reference class (eilev.model.v2, installed unmodified into baseline/_ref by
    baseline/install_ref.py) through model(**batch) + loss.backward(): fp32 on all host cores
    (device="cpu") or bf16 + HF SDPA on the B200 (device="cuda", the library bar).  Falls back to the
    oracle port on the CPU when the reference is not installed."""
    from baseline import install_ref

    threads = os.cpu_count() or 1
    batch = synthetic_batch(1, clips=clips)
    seq = int(batch["input_ids"].shape[1])
    if not install_ref.available():
        if device != "cpu":
            raise RuntimeError("baseline/_ref is not installed (run baseline/install_ref.py where /root/reference exists)")
        return run_cpu_baseline(cfg, clips, steps, warmup)
    from baseline import reference_arm

    dt, loss, build_s = reference_arm.run(cfg, batch, device, steps, warmup, threads)
    what = ("fp32, %d host threads" % threads) if device == "cpu" else "bf16 weights, HF SDPA attention, one B200"
    return dict(value=clips / dt, unit="clips/s", cores=threads if device == "cpu" else 0, kind="reference",
                sample=f"{steps} x fwd+bwd of {clips} clip(s) x {FRAMES} frames (L={seq}) after {warmup} warm-up, full-size "
                       f"random-init eilev-blip2-opt-2.7b, the reference's own VideoBlipForConditionalGeneration "
                       f"(baseline/_ref, unmodified) via model(**batch) + loss.backward(), {what}",
                ms_per_step=dt * 1e3, loss=loss, build_s=build_s, seq_len=seq, clips=clips)


def reference_arm(args) -> None:
    """bench.py --impl reference: ONE timed fwd+bwd of the TRUE benchmark datapoint (17 clips, L = 976)
    by the real reference class on the host cores — about 2-3 minutes; --steps/--warmup are not
    multiplied into it (20 steps would take the better part of an hour), and the line says so.
    --device cuda times the same class on the B200 instead (bf16, SDPA): the library bar."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = full_config(args.dropout)
    clips = args.cpu_clips if args.cpu_clips > 0 else CLIPS
    if args.device == "cuda":
        steps, warmup = max(1, min(args.steps, 10)), max(2, min(args.warmup, 3))
    else:
        steps, warmup = 1, 0
    res = run_reference(cfg, clips, steps, warmup, args.device)
    name = workload_config(args.gpus, res.get("seq_len", 976), None, args.dropout)
    name["workload"] = (f"eilev-blip2-opt-2.7b 16-ctx x 8-frame fwd+bwd bs=1 ({clips} clips, L={res.get('seq_len', 976)}): "
                        f"{res['sample']}; optimizer step not included")
    name["steps_requested"] = args.steps
    line = {
        "impl": "reference", "metric": "clips/sec fwd+bwd (8-frame x 17-ctx)", "value": res["value"],
        "unit": "clips/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.device == "cpu" else "bf16", "data": "synthetic", "device": args.device,
        "config": name,
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def build_gpu_model(cfg, device):
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from eilev_b200.train import freeze_for_recipe

    torch.manual_seed(1234)
    with torch.device(device):
        model = VideoBlipForConditionalGeneration(cfg)
    g = torch.Generator(device=device).manual_seed(1234)
    with torch.no_grad():
        for name, p in model.named_parameters():
            ln = name.lower()
            if "layernorm" in ln or "layer_norm" in ln:
                p.fill_(1.0) if name.endswith("weight") else p.zero_()
            else:
                p.normal_(0.0, 0.02, generator=g)
    model = model.to(torch.bfloat16)  # bf16-resident frozen towers
    freeze_for_recipe(model)
    for p in model.parameters():
        if p.requires_grad:
            p.data = p.data.float()  # f32 master copies of the trainable 107 M
    return model.train()


class GemmProfiler:
    """Per-launch CUDA-event timing of the tcgen05 GEMM (the dominant kernel), used for ONE
    instrumented step after the timed region."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        from eilev_b200 import ops

        self.ops, self.orig = ops, ops.gemm
        prof = self

        def timed(a, w, *args, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = prof.orig(a, w, *args, **kw)
            e.record()
            prof.records.append((a.shape[0], w.shape[0], a.shape[1], s, e, prof.ops.gemm_uses_tcgen05(a, w, out)))
            return out

        ops.gemm = timed
        for mod in ("vision", "qformer", "opt"):
            getattr(__import__(f"eilev_b200.engine.{mod}", fromlist=["ops"]), "ops").gemm = timed
        return self

    def __exit__(self, *exc):
        self.ops.gemm = self.orig
        return False

    def summary(self):
        torch.cuda.synchronize()
        flops = ms = 0.0
        n = 0
        self.shapes = []
        for m, nn_, k, s, e, tc in self.records:
            if tc and m >= 4096:  # the ViT-sized launches that dominate the step
                flops += 2.0 * m * nn_ * k
                ms += s.elapsed_time(e)
                n += 1
                self.shapes.append((m, nn_, k))
        return flops, ms, n


def measured_traffic(shapes):
    """DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one
    ``ncu --set full`` capture per GEMM shape (committed summary profiles/r*_ncu_gemm_traffic.json, written
    by scripts/ncu_summary.py --traffic), averaged over the launches `roofline.achieved` averages over.
    Returns (bytes per launch | None, algorithmic bytes per launch, source file | None)."""
    algo = [2.0 * (m * k + n * k + m * n) for m, n, k in shapes]  # bf16 A + B read once, C written once
    algo_mean = sum(algo) / len(algo) if algo else None
    files = sorted((ROOT / "profiles").glob("r*_ncu_gemm_traffic.json"))
    if not files or not shapes:
        return None, algo_mean, None
    table = {tuple(r["mnk"]): float(r["dram_bytes"]) for r in json.loads(files[-1].read_text())["launches"]}
    got = [table.get((m, n, k)) for m, n, k in shapes]
    if any(g is None for g in got):
        # shapes without a capture count with their algorithmic bytes scaled by the mean measured ratio
        known = [(g, a) for g, a in zip(got, algo) if g is not None]
        if not known:
            return None, algo_mean, files[-1].name
        ratio = sum(g for g, _ in known) / sum(a for _, a in known)
        got = [g if g is not None else a * ratio for g, a in zip(got, algo)]
    return sum(got) / len(got), algo_mean, files[-1].name


def measure_decode(model, device, batch: int = 1):
    """BASELINE configs[4]: greedy generate() after the 16-context prompt (17 clips, L = 958).
    tok/s of the CUDA-graphed decode loop after the prefill, CUDA events over 64 steps (vision
    tower and prefill excluded); bytes per token = 5.293 GB of bf16 weights + the KV pages read."""
    was_training = model.training
    model.eval()
    one = synthetic_batch(7)
    n = int(one["attention_mask"].sum()) - TARGET_TOKENS
    ids = one["input_ids"][:, :n].repeat(batch, 1).to(device)
    vm = one["video_input_mask"][:, :n].repeat(batch, 1).to(device)
    px = one["pixel_values"].to(device)
    if batch > 1:
        px = px.repeat(batch, 1, 1, 1, 1)

    from eilev_b200.engine import opt as E_opt

    steps = 64
    with torch.no_grad():
        feats, _, _ = model._video_features(px, False, train=False)
        lm = model.language_model
        logits, state = E_opt.opt_prefill(lm, lm._pack, ids, torch.ones_like(ids), vm.bool(), feats, steps + 8)
        graph = E_opt.DecodeGraph(lm, lm._pack, state, batch, device)
        tok = logits.argmax(-1)
        for _ in range(3):  # warm-up replays
            tok = graph.step(tok).argmax(-1)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):  # the greedy loop of generate(): graph replay + device argmax
            tok = graph.step(tok).argmax(-1)
        e.record()
        torch.cuda.synchronize()
    per_tok = s.elapsed_time(e) * 1e-3 / steps
    model.train(was_training)
    return {"metric": "decode tok/s (greedy, 16-ctx prompt, batch %d)" % batch, "value": batch / per_tok,
            "unit": "tok/s", "ms_per_token": per_tok * 1e3, "prompt_len": n, "batch": batch, "steps": steps,
            "bytes_per_token": 5.293e9 + 327680.0 * (n + 3 + steps / 2) * batch}


def measure_generate(model, device, batch: int = 1, new_tokens: int = 64):
    """Decode tok/s THROUGH THE PUBLIC API: wall-clock of ``model.generate(max_new_tokens=1 + n,
    min_new_tokens=1 + n)`` minus ``model.generate(max_new_tokens=1, min_new_tokens=1)`` (vision tower +
    Q-Former + prefill + first token), inputs copied from pinned host memory and the generated ids read
    back to the host inside each call (eilev/model/v2.py:254-324; samples/eilev_generate_action_narration.py:60-75)."""
    was_training = model.training
    model.eval()
    one = synthetic_batch(7)
    n = int(one["attention_mask"].sum()) - TARGET_TOKENS
    host = dict(input_ids=one["input_ids"][:, :n].repeat(batch, 1).pin_memory(),
                video_input_mask=one["video_input_mask"][:, :n].repeat(batch, 1).pin_memory(),
                attention_mask=torch.ones(batch, n, dtype=torch.long).pin_memory(),
                pixel_values=one["pixel_values"].repeat(batch, 1, 1, 1, 1).pin_memory())

    def call(k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev = {key: v.to(device, non_blocking=True) for key, v in host.items()}
        ids = model.generate(**dev, max_new_tokens=k, min_new_tokens=k, do_sample=False)
        ids = ids.cpu()
        dt = time.perf_counter() - t0
        assert ids.shape == (batch, k), ids.shape
        return dt

    call(1 + new_tokens)  # warm-up: packs, graph capture paths, allocator
    t_long = min(call(1 + new_tokens) for _ in range(3))
    t_short = min(call(1) for _ in range(3))
    model.train(was_training)
    per_tok = (t_long - t_short) / new_tokens
    return {"value": batch / per_tok, "unit": "tok/s", "ms_per_token": per_tok * 1e3,
            "generate_ms": t_long * 1e3, "prefill_call_ms": t_short * 1e3, "new_tokens": new_tokens,
            "h2d_bytes_per_call": sum(v.numel() * v.element_size() for v in host.values()),
            "d2h_bytes_per_call": batch * (1 + new_tokens) * 8,
            "how": "model.generate(max_new_tokens=65) minus model.generate(max_new_tokens=1), host wall clock, best of 3"}


def measure_decode_t5(model, device, steps: int = 32):
    """BASELINE configs[3] inference side: greedy generate() with the flan-t5-xl LM after the
    16-context prompt (encoder + cross-attention K|V once, excluded; the decoder re-runs its
    prefix every step).  Bytes per token = decoder + head weights (bf16) + the cross K|V read."""
    from eilev_b200.model.generation import _T5Stepper

    was_training = model.training
    model.eval()
    one = synthetic_batch(7, lm="t5")
    ids, am = one["input_ids"].to(device), one["attention_mask"].to(device)
    vm, px = one["video_input_mask"].to(device), one["pixel_values"].to(device)
    cfg = model.config.text_config
    inner = cfg.num_heads * cfg.d_kv
    per_layer = 4 * cfg.d_model * inner + 2 * cfg.d_model * inner + 3 * cfg.d_model * cfg.d_ff
    weight_bytes = 2.0 * (cfg.num_decoder_layers * per_layer + cfg.vocab_size * cfg.d_model)
    kv_bytes = 2.0 * cfg.num_decoder_layers * 2 * ids.shape[1] * inner
    with torch.no_grad():
        feats, _, _ = model._video_features(px, False, train=False)
        stepper = _T5Stepper(model.language_model)
        logits = stepper.prefill(ids, am, vm.bool(), feats, steps + 4)
        stepper.graph(1, device)  # the CUDA-graphed fixed-length step generate() uses
        tok = logits.argmax(-1)
        for _ in range(3):
            tok = stepper.step(tok).argmax(-1)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            tok = stepper.step(tok).argmax(-1)
        e.record()
        torch.cuda.synchronize()
    per_tok = s.elapsed_time(e) * 1e-3 / steps
    model.train(was_training)
    return {"metric": "decode tok/s (greedy, flan-t5-xl, 16-ctx prompt, batch 1)", "value": 1.0 / per_tok,
            "unit": "tok/s", "ms_per_token": per_tok * 1e3, "prompt_len": int(ids.shape[1]), "batch": 1,
            "steps": steps, "bytes_per_token": weight_bytes + kv_bytes,
            "note": "KV-cached decoder step on the weight-streaming GEMV kernels, one CUDA graph"}


def gpu_arm(args) -> None:
    import torch.distributed as dist

    from eilev_b200 import _lib
    from eilev_b200.train import DataParallelTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.lib()  # fail loudly if the extension is missing

    cfg = full_config(args.dropout, args.lm)
    model = build_gpu_model(cfg, device)
    decode = decode8 = None
    if rank == 0 and not args.no_decode and not args.profile and args.lm == "opt":
        # independent workloads (BASELINE configs[4], batch sweep ends), measured before the training loop
        decode = measure_decode(model, device)
        decode["e2e"] = measure_generate(model, device)
        decode8 = measure_decode(model, device, batch=8)
        decode8["e2e"] = measure_generate(model, device, batch=8)
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_decode and not args.profile and args.lm == "t5":
        decode = measure_decode_t5(model, device)
    trainer = DataParallelTrainer(model, lr=1e-5, weight_decay=0.05, max_grad_norm=1.0,
                                  grad_accum=GRAD_ACCUM)
    if world > 1:  # NCCL communicator / NVLink connection set-up happens on the first collective
        for _ in range(2):
            dist.all_reduce(trainer.flat.grads)
        torch.cuda.synchronize()
    host = synthetic_batch(1000 + rank, lm=args.lm)
    if args.u8_frames:  # opt-in: decoded uint8 frames, normalised inside the patch gather (not the headline config)
        g8 = torch.Generator().manual_seed(2000 + rank)
        host["pixel_values"] = torch.randint(0, 256, tuple(host["pixel_values"].shape), dtype=torch.uint8, generator=g8)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(device) for k, v in host.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident():
        trainer.micro_step(resident)

    def step_e2e():
        batch = {k: v.to(device, non_blocking=True) for k, v in pinned.items()}
        loss = trainer.micro_step(batch)
        return float(loss)  # device -> host read of the step's result

    if not args.no_graph:
        trainer.capture_graph(resident)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    if args.profile:  # ncu --profile-from-start off: exactly one (warm, no re-pack) micro-step is captured
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # untimed, up to the next accumulation boundary: the first optimizer step (global norm, clip, fused AdamW and the
    # torch glue between them) pays its one-time launch costs here; the timed region still runs one per grad_accum
    while trainer.micro % trainer.grad_accum != 0:
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    calls0 = _lib.launch_count()
    ms = timed_region(step_resident, args.steps)
    launches = _lib.launch_count() - calls0
    if getattr(trainer, "_graph", None) is not None:  # replayed launches are not re-counted
        launches += trainer.launches_per_graph * args.steps
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed_region(step_e2e, args.steps)

    # roofline of the dominant kernel: instrumented (eager, un-graphed) steps.  Three of them: a single eager step
    # read 1145-1280 TFLOP/s from run to run on the same code (the SM clock under the power cap moves with the idle
    # gaps of eager launching); the ratio of the sums is what is reported, the launch count is per step.
    trainer._graph = None
    trainer.micro_step(resident)  # warm the caching allocator of this stream (no cudaMalloc inside the timed launches)
    with GemmProfiler() as prof:
        for _ in range(3):
            trainer.micro_step(resident)
    g_flops, g_ms, g_n = prof.summary()
    g_n //= 3
    prof.shapes = prof.shapes[:g_n]

    if rank == 0:
        peaks = {}
        pfile = ROOT / "MEASURED_PEAKS.json"
        if pfile.exists():
            peaks = json.loads(pfile.read_text())
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        traffic, traffic_algo, traffic_src = measured_traffic(prof.shapes)
        clips_per_s = world * CLIPS * args.steps / (ms * 1e-3)
        e2e_clips = world * CLIPS * args.steps / (ms_e2e * 1e-3)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        line = {
            "metric": "clips/sec fwd+bwd (8-frame x 17-ctx)", "value": clips_per_s, "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(world, int(host["input_ids"].shape[1]), not args.no_graph, args.dropout,
                                           args.lm),
                           **({"frames": "uint8 pixel_values, rescale + normalize fused into the patch gather"}
                              if args.u8_frames else {})),
            "clocks": clocks,
            "e2e": {"value": e2e_clips, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_algorithmic": traffic_algo, "traffic_source": traffic_src,
                         "kernel": "gemm_tcgen05_2cta_kernel (ViT / cross-K|V launches with M>=4096)",
                         "launches": g_n, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"
                         if peaks else "fallback 1400"},
            "step_flops_frac": (FLOPS_PER_DATAPOINT if args.lm == "opt" else t5_flops_per_datapoint())
            * args.steps / (ms * 1e-3) / (peak * 1e12) if peak else None,
        }
        if decode is not None:
            hbm = peaks.get("hbm_gbs", 6650.0)
            decode["roofline"] = {"bound": "hbm", "achieved": decode["bytes_per_token"] / (decode["ms_per_token"] * 1e-3) / 1e9,
                                  "peak": hbm, "unit": "GB/s",
                                  "frac": decode["bytes_per_token"] / (decode["ms_per_token"] * 1e-3) / 1e9 / hbm}
            line["decode"] = decode
            if decode8 is not None:
                decode8["roofline"] = {"bound": "hbm", "achieved": decode8["bytes_per_token"] / (decode8["ms_per_token"] * 1e-3) / 1e9,
                                       "peak": hbm, "unit": "GB/s",
                                       "frac": decode8["bytes_per_token"] / (decode8["ms_per_token"] * 1e-3) / 1e9 / hbm}
                line["decode_batch8"] = decode8
        if not args.no_library_bar and world == 1 and args.lm == "opt":
            # the reference's own class on this B200 (bf16, HF SDPA): the bar the kernels have to beat
            del trainer
            model._pack.clear()
            torch.cuda.empty_cache()
            try:
                res = run_reference(cfg, CLIPS, steps=5, warmup=2, device="cuda")
                line["library_bar"] = {"value": res["value"], "unit": "clips/s", "ms_per_step": res["ms_per_step"],
                                       "kind": res["kind"], "sample": res["sample"],
                                       "speedup_device": clips_per_s / res["value"], "speedup_e2e": e2e_clips / res["value"]}
            except Exception as exc:  # the reference is optional on the box; say why it is missing
                line["library_bar"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()
        if not args.no_cpu_baseline and world == 1 and args.lm == "opt":
            # bounded sample of the same workload (2 of the 17 clips, L = 120) by the real reference class on
            # the host cores; `bench.py --impl reference` times the whole 17-clip datapoint
            res = run_reference(cfg, clips=args.cpu_clips if args.cpu_clips > 0 else 2, steps=2, warmup=0, device="cpu")
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-clips", type=int, default=0,
                    help="clips per CPU step: default 17 (the true datapoint) for --impl reference, 2 for the "
                         "bounded cpu_baseline sample inside the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-bar", action="store_true", help="skip timing the reference class on the GPU")
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu = the reference arm (fp32, host cores); cuda = the library "
                         "bar (the reference class in bf16 with HF SDPA on the B200)")
    ap.add_argument("--profile", action="store_true", help="run one profiler-bracketed step and exit")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the CUDA graph")
    ap.add_argument("--no-decode", action="store_true", help="skip the decode tok/s measurement")
    ap.add_argument("--lm", default="opt", choices=["opt", "t5"],
                    help="language model of the workload: opt = eilev-blip2-opt-2.7b (the headline), "
                         "t5 = eilev-blip2-flan-t5-xl (BASELINE configs[3], fwd+bwd only)")
    ap.add_argument("--u8-frames", action="store_true",
                    help="feed decoded uint8 frames (rescale + normalize fused into the patch gather; 20.5 MB instead "
                         "of 82 MB host->device per datapoint) instead of the reference's processed fp32 pixel_values")
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="dropout of the training step (0.1 = the reference recipe; 0 = parity configuration)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
