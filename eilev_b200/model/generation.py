"""Native decoding loops over the paged-KV OPT engine.

Mirrors what the reference reaches through ``language_model.generate(inputs_embeds=...,
attention_mask=..., **generate_kwargs)`` (eilev/model/v2.py:318-322 ->
HF:generation/utils.py ``_sample`` / ``_beam_search``) for the keyword arguments its callers
use (samples/eilev_generate_action_narration.py:60-75, demo/eilev_demo.py:51-67,
scripts/general/generate_narration_texts.py:116-123): ``max_new_tokens``, ``min_new_tokens``,
``num_beams``, ``do_sample`` (+ ``temperature`` / ``top_k`` / ``top_p``), ``length_penalty``,
``repetition_penalty``, ``eos_token_id``, ``pad_token_id``, ``early_stopping``.

As in the reference (decoder-only LM fed with ``inputs_embeds``) only the NEW tokens are
returned; finished rows are padded with ``pad_token_id``.  Prompts are expected
left-padded when batched (generate_narration_texts.py:229-230).  The per-step logits come
from the CUDA engine; the token bookkeeping on (B*beams,)-sized tensors is host-side torch.
"""
from __future__ import annotations

import torch

from ..engine import opt as E_opt
from ..engine import t5 as E_t5

class _OptStepper:
    """Decoder-only LM: prompt prefill into the paged KV cache, then one token per step."""

    start_token = None

    def __init__(self, lm, slot: int = 0) -> None:
        self.lm = lm
        self.slot = slot  # generation-state slot: steppers alive at the same time must not share KV pools

    def prefill(self, input_ids, attention_mask, video_mask, feats, max_new):
        logits, self.state = E_opt.opt_prefill(self.lm, self.lm._pack, input_ids, attention_mask, video_mask,
                                               feats, max_new, reuse_slot=self.slot)
        self.status = self.state["status"]
        return logits

    def graph(self, rows, dev):
        return E_opt.decode_graph_for(self.lm, self.lm._pack, self.state, rows, dev)

    def step(self, tokens):
        return E_opt.opt_decode_step(self.lm, self.lm._pack, tokens, self.state)

    def reorder(self, src) -> None:
        self.state["kv"].reorder(src)
        for key in ("ctx_len", "first_valid", "n_valid"):
            self.state[key].copy_(self.state[key][src])  # in place: the decode program holds these pointers


class _T5Stepper:
    """Encoder-decoder LM (flan-T5): the encoder and the cross-attention K|V run once.  Up to 16
    rows (batch x beams) decode token by token on the weight-streaming kernels with a paged
    self-attention cache (engine/t5.py::t5_decode_step), one CUDA graph per generate call;
    larger batches re-run the decoder over the prefix (t5_decode_logits)."""

    def __init__(self, lm) -> None:
        self.lm = lm
        self.start_token = int(lm.config.decoder_start_token_id)
        self._graph = None
        self.cached = False

    def prefill(self, input_ids, attention_mask, video_mask, feats, max_new):
        lm = self.lm
        self.enc = E_t5.t5_encode(lm, lm._pack, input_ids, attention_mask, video_mask, feats)
        self.status = self.enc["status"]
        b, dev = input_ids.shape[0], input_ids.device
        start = torch.full((b,), self.start_token, dtype=torch.long, device=dev)
        self.cached = b <= 16 and (lm.config.num_heads * lm.config.d_kv) % 8 == 0
        if self.cached:
            self.st = E_t5.t5_decode_init(lm, lm._pack, self.enc, max_new)
            return E_t5.t5_decode_step(lm, lm._pack, self.enc, self.st, start)
        self.prefix = start.view(b, 1)
        return E_t5.t5_decode_logits(lm, lm._pack, self.enc, self.prefix)

    def graph(self, rows, dev):
        """Captures the cached decoder step; returns an object with .step(tokens)."""
        if not self.cached:
            return None
        lm = self.lm
        self._tokens = torch.zeros(rows, dtype=torch.long, device=dev)
        snap = self.st["ctx_len"].clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up: kernel attributes, allocator
            E_t5.t5_decode_step(lm, lm._pack, self.enc, self.st, self._tokens)
        torch.cuda.current_stream().wait_stream(side)
        self.st["ctx_len"].copy_(snap)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._logits = E_t5.t5_decode_step(lm, lm._pack, self.enc, self.st, self._tokens)
        self.st["ctx_len"].copy_(snap)
        return self

    def step(self, tokens):
        lm = self.lm
        if self._graph is not None:
            self._tokens.copy_(tokens.view(-1))
            self._graph.replay()
            return self._logits
        if self.cached:
            return E_t5.t5_decode_step(lm, lm._pack, self.enc, self.st, tokens.view(-1))
        self.prefix = torch.cat([self.prefix, tokens.view(-1, 1)], dim=1)
        return E_t5.t5_decode_logits(lm, lm._pack, self.enc, self.prefix)

    def reorder(self, src) -> None:
        # beams of one prompt share its encoder rows and all rows have the same length
        if self.cached:
            self.st["kv"].reorder(src)
        else:
            self.prefix = self.prefix[src]


class _GroupedStepper:
    """More rows (batch x beams) than one weight-streaming decode step takes (16): the rows are dealt
    to groups of whole beam sets, each group a stepper of its own (own paged KV cache, own CUDA
    graph); a step streams the weights once per group.  Beam search permutes rows only inside one
    prompt's beams, so ``reorder`` never crosses a group.  Token bookkeeping only — the arithmetic
    stays in the per-group steppers."""

    def __init__(self, make, rows_per_group: int) -> None:
        assert rows_per_group >= 1
        self.make, self.rows_per_group = make, int(rows_per_group)
        self.start_token = None
        self.groups: list = []
        self.bounds: list[tuple[int, int]] = []
        self.graphs: list | None = None

    def prefill(self, input_ids, attention_mask, video_mask, feats, max_new):
        rows = input_ids.shape[0]
        self.bounds = [(r0, min(rows, r0 + self.rows_per_group)) for r0 in range(0, rows, self.rows_per_group)]
        # video features are spliced in row-major mask order: a group's rows own a contiguous slice
        offsets = [0] * (rows + 1)
        if feats is not None:
            counts = video_mask.sum(dim=1).tolist()
            for r, c in enumerate(counts):
                offsets[r + 1] = offsets[r] + int(c)
        logits, statuses = [], []
        self.groups = []
        for r0, r1 in self.bounds:
            st = self.make()
            self.groups.append(st)
            logits.append(st.prefill(input_ids[r0:r1], attention_mask[r0:r1],
                                     None if video_mask is None else video_mask[r0:r1],
                                     None if feats is None else feats[offsets[r0]:offsets[r1]], max_new))
            statuses.append(getattr(st, "status", None))
        self.start_token = self.groups[0].start_token
        live = [x for x in statuses if x is not None]
        self.status = torch.stack(live).sum(dim=0).to(live[0].dtype) if live else None  # (mismatch flag, slot count)
        return torch.cat(logits, dim=0)

    def graph(self, rows, dev):
        self.graphs = [st.graph(r1 - r0, dev) for st, (r0, r1) in zip(self.groups, self.bounds)]
        return self

    def step(self, tokens):
        tokens = tokens.view(-1)
        out = []
        for i, (st, (r0, r1)) in enumerate(zip(self.groups, self.bounds)):
            g = self.graphs[i] if self.graphs is not None else None
            out.append((g if g is not None else st).step(tokens[r0:r1]))
        return torch.cat(out, dim=0)  # copies: a graphed group returns its static logits buffer

    def reorder(self, src) -> None:
        for st, (r0, r1) in zip(self.groups, self.bounds):
            local = src[r0:r1] - r0
            if not bool(((local >= 0) & (local < r1 - r0)).all()):
                raise RuntimeError("beam reorder crosses a decode group")
            st.reorder(local)


MAX_DECODE_ROWS = 16  # rows one weight-streaming decode step takes (engine/opt.py::opt_decode_step)

_UNSUPPORTED = ("penalty_alpha", "num_beam_groups", "diversity_penalty", "constraints",
                "force_words_ids", "assistant_model", "prompt_lookup_num_tokens")


def _as_list(x):
    if x is None:
        return []
    if isinstance(x, (list, tuple)):
        return list(x)
    if isinstance(x, torch.Tensor):
        return x.flatten().tolist()
    return [int(x)]


def _process_logits(logits, generated, step, *, min_new_tokens, eos_ids, repetition_penalty, start_token=None):
    """Logit processors in HF order: repetition penalty, then min-length EOS suppression.
    start_token: an encoder-decoder LM's decoder_start_token — HF hands the processors the
    decoder input ids, so the start token counts as "already generated" for the penalty."""
    if start_token is not None:
        generated = torch.cat([torch.full((generated.shape[0], 1), start_token, dtype=generated.dtype,
                                          device=generated.device), generated], dim=1)
    if repetition_penalty and repetition_penalty != 1.0 and generated.shape[1] > 0:
        score = torch.gather(logits, 1, generated)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        logits = logits.scatter(1, generated, score)
    if eos_ids and step < min_new_tokens:
        logits = logits.clone()
        logits[:, eos_ids] = float("-inf")
    return logits


def _warp(logits, temperature, top_k, top_p):
    if temperature and temperature != 1.0:
        logits = logits / temperature
    if top_k and top_k > 0:
        k = min(int(top_k), logits.shape[-1])
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p is not None and top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=False)
        cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= (1 - top_p)
        remove[..., -1:] = False
        remove = remove.scatter(1, sorted_idx, remove)
        logits = logits.masked_fill(remove, float("-inf"))
    return logits


@torch.no_grad()
def generate(model, input_ids, attention_mask, video_mask, video_features, **kw) -> torch.Tensor:
    for k in _UNSUPPORTED:
        if kw.get(k):
            raise NotImplementedError(f"generate({k}=...) is not supported by the native decoder")
    gcfg = kw.pop("generation_config", None)

    def opt(name, default):
        if name in kw and kw[name] is not None:
            return kw[name]
        if gcfg is not None and getattr(gcfg, name, None) is not None:
            return getattr(gcfg, name)
        mg = getattr(model, "generation_config", None)
        if mg is not None and getattr(mg, name, None) is not None and name not in ("max_length",):
            return getattr(mg, name)
        return default

    tcfg = model.config.text_config
    max_new = kw.get("max_new_tokens")
    if max_new is None and gcfg is not None:
        max_new = getattr(gcfg, "max_new_tokens", None)
    if max_new is None:
        max_length = kw.get("max_length") or 20
        max_new = max(int(max_length), 1)  # prompt is embeddings: length counts new tokens only
    max_new = int(max_new)
    min_new = int(opt("min_new_tokens", 0) or 0)
    num_beams = int(opt("num_beams", 1) or 1)
    do_sample = bool(opt("do_sample", False))
    eos_ids = _as_list(kw["eos_token_id"] if "eos_token_id" in kw else
                       (getattr(gcfg, "eos_token_id", None) if gcfg is not None else tcfg.eos_token_id))
    pad_id = opt("pad_token_id", tcfg.pad_token_id)
    if pad_id is None:
        pad_id = eos_ids[0] if eos_ids else 0
    rep = float(opt("repetition_penalty", 1.0) or 1.0)
    temperature = float(opt("temperature", 1.0) or 1.0)
    top_k = opt("top_k", 50 if do_sample else 0)
    top_p = opt("top_p", 1.0)
    length_penalty = float(opt("length_penalty", 1.0))
    early_stopping = opt("early_stopping", False)
    if int(opt("num_return_sequences", 1) or 1) != 1:
        raise NotImplementedError("num_return_sequences > 1 is not supported")

    lm = model.language_model
    b = input_ids.shape[0]
    dev = input_ids.device
    if num_beams > 1:
        # expand every prompt to num_beams rows (HF _expand_inputs_for_generation)
        rep_idx = torch.arange(b, device=dev).repeat_interleave(num_beams)
        n_per = None
        if video_features is not None:
            # features are spliced in row-major mask order: replicate each row's features
            counts = video_mask.sum(dim=1).tolist()
            chunks, start = [], 0
            for c in counts:
                chunks += [video_features[start:start + c]] * num_beams
                start += c
            video_features = torch.cat(chunks, dim=0) if chunks else video_features
            del n_per
        input_ids = input_ids[rep_idx]
        attention_mask = attention_mask[rep_idx]
        video_mask = video_mask[rep_idx] if video_mask is not None else None

    stepper = kw.pop("_stepper", None)  # test seam: the decoding loops over an injected LM stepper
    if stepper is None:
        if not model.config.use_decoder_only_language_model:
            stepper = _T5Stepper(lm)
        elif input_ids.shape[0] <= MAX_DECODE_ROWS:
            stepper = _OptStepper(lm)
        else:
            if num_beams > MAX_DECODE_ROWS:
                raise NotImplementedError(f"num_beams > {MAX_DECODE_ROWS} is not supported")
            slots = iter(range(1, 1 << 20))
            stepper = _GroupedStepper(lambda: _OptStepper(lm, next(slots)), MAX_DECODE_ROWS // num_beams * num_beams)
    logits = stepper.prefill(input_ids, attention_mask, video_mask, video_features, max_new)
    model._last_splice_status = stepper.status

    def finish(tokens):
        """HF returns only the new tokens for a decoder-only LM fed with embeddings, and
        [decoder_start_token] + new tokens for an encoder-decoder LM."""
        if stepper.start_token is None:
            return tokens
        return torch.cat([torch.full((tokens.shape[0], 1), stepper.start_token, dtype=torch.long, device=dev),
                          tokens], dim=1)

    if num_beams > 1:
        # the decode step of the beam rows as one CUDA graph (the reorder between steps permutes the state's page
        # tables / counters in place, which is what the captured kernels read)
        beam_rows = b * num_beams
        graph_ok = bool(kw.get("use_cuda_graph", max_new >= 8)) and max_new > 1 and (
            beam_rows <= MAX_DECODE_ROWS or isinstance(stepper, _GroupedStepper))
        bgraph = stepper.graph(beam_rows, dev) if graph_ok else None
        return finish(_beam_search(stepper, logits, b, num_beams, max_new, min_new, eos_ids, pad_id, rep,
                                   length_penalty, early_stopping, do_sample, temperature, top_k, top_p,
                                   graph=bgraph))

    rows = input_ids.shape[0]
    use_graph = bool(kw.get("use_cuda_graph", max_new >= 8)) and (rows <= MAX_DECODE_ROWS or isinstance(stepper, _GroupedStepper))
    dgraph = stepper.graph(rows, dev) if use_graph and max_new > 1 else None
    # Token bookkeeping stays on the device: the new tokens go into a preallocated (rows, max_new)
    # buffer and the "every row has emitted EOS" test (HF's stopping criterion) is read back only
    # every `poll` steps, so the host keeps queueing decode steps instead of draining the stream
    # once per token.  Rows that finished emit pad_id, so the columns decoded past the stopping
    # step are cut off below and the returned ids equal the step-by-step loop's.
    poll = int(kw.get("eos_poll_interval", 8))
    if (dgraph is not None and isinstance(stepper, _OptStepper) and not do_sample and (not rep or rep == 1.0)
            and stepper.start_token is None and bool(kw.get("device_bookkeeping", True))):
        # plain greedy search on the decoder-only LM: the whole iteration (EOS suppression, argmax, pad for
        # finished rows, the write into the output buffer, the alive count, the next decode step) is one graph
        # replay with a device-side position counter (engine/opt.py::DecodeGraph.greedy_*)
        lm = stepper.lm
        dgraph.greedy_begin(lm, lm._pack, stepper.state, logits, max_new, min_new, eos_ids, int(pad_id))
        n_done = max_new
        for step in range(max_new):
            if step + 1 < max_new:
                dgraph.greedy_iteration()
            else:
                dgraph.greedy_last()
            if eos_ids and step >= min_new - 1 and ((step + 1) % poll == 0 or step + 1 == max_new):
                dead = (dgraph.g_alive[: step + 1] == 0).nonzero()
                if dead.numel():  # one device -> host read per `poll` tokens
                    n_done = int(dead[0]) + 1
                    break
        return finish(dgraph.g_out[:, :n_done].clone())
    out = torch.full((rows, max_new), int(pad_id), dtype=torch.long, device=dev)
    pad_t = torch.full((rows,), int(pad_id), dtype=torch.long, device=dev)
    unfinished = torch.ones(rows, dtype=torch.bool, device=dev)
    eos_t = torch.tensor(eos_ids, device=dev, dtype=torch.long) if eos_ids else None
    alive = torch.ones(max_new, dtype=torch.int32, device=dev) if eos_t is not None else None
    n_done = max_new
    for step in range(max_new):
        scores = _process_logits(logits, out[:, :step], step, min_new_tokens=min_new, eos_ids=eos_ids,
                                 repetition_penalty=rep, start_token=stepper.start_token)
        if do_sample:
            probs = _warp(scores, temperature, top_k, top_p).softmax(dim=-1)
            nxt = torch.multinomial(probs, 1).squeeze(1)
        else:
            nxt = scores.argmax(dim=-1)
        nxt = torch.where(unfinished, nxt, pad_t)
        out[:, step] = nxt
        if eos_t is not None:
            unfinished = unfinished & ~torch.isin(nxt, eos_t)
            alive[step] = unfinished.sum()
            if step >= min_new - 1 and ((step + 1) % poll == 0 or step + 1 == max_new):
                dead = (alive[: step + 1] == 0).nonzero()
                if dead.numel():  # one device -> host read per `poll` tokens
                    n_done = int(dead[0]) + 1
                    break
        if step + 1 < max_new:
            logits = dgraph.step(nxt) if dgraph is not None else stepper.step(nxt)
    return finish(out[:, :n_done])


def _beam_search(stepper, logits, batch, nb, max_new, min_new, eos_ids, pad_id, rep,
                 length_penalty, early_stopping, do_sample, temperature, top_k, top_p, graph=None):
    """Standard beam search with HF's scoring: hypotheses are ranked by
    sum_logprobs / generated_len**length_penalty (BeamHypotheses.add)."""
    dev = logits.device
    vocab = logits.shape[-1]
    beam_scores = torch.zeros((batch, nb), device=dev)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    generated = torch.empty((batch * nb, 0), dtype=torch.long, device=dev)
    finished: list[list[tuple[float, torch.Tensor]]] = [[] for _ in range(batch)]
    done = [False] * batch
    eos_set = set(eos_ids)

    def worst(i):
        return min(h[0] for h in finished[i])

    for step in range(max_new):
        # HF beam search normalises FIRST and runs the logits processors on the log-probabilities
        # (generation/utils.py::_beam_search step b): a suppressed EOS does not renormalise the rest
        logp = _process_logits(torch.log_softmax(logits.float(), dim=-1), generated, step, min_new_tokens=min_new,
                               eos_ids=eos_ids, repetition_penalty=rep, start_token=stepper.start_token)
        if do_sample:
            logp = torch.log_softmax(_warp(logp, temperature, top_k, top_p), dim=-1)
        cand = (logp + beam_scores[:, None]).view(batch, nb * vocab)
        if do_sample:
            pick = torch.multinomial(cand.softmax(dim=-1), 2 * nb)
            top_s = torch.gather(cand, 1, pick)
            top_s, order = top_s.sort(dim=1, descending=True)
            top_i = torch.gather(pick, 1, order)
        else:
            top_s, top_i = torch.topk(cand, 2 * nb, dim=1)
        top_s_l, top_i_l = top_s.tolist(), top_i.tolist()
        next_scores = torch.zeros((batch, nb), device=dev)
        next_tokens = torch.full((batch, nb), pad_id, dtype=torch.long, device=dev)
        next_src = torch.zeros((batch, nb), dtype=torch.long, device=dev)
        gen_len = step + 1
        for i in range(batch):
            base = i * nb
            if done[i]:
                next_src[i] = base
                continue
            keep = []
            for rank, (s, idx) in enumerate(zip(top_s_l[i], top_i_l[i])):
                src, tok = idx // vocab, idx % vocab
                if tok in eos_set:
                    if rank >= nb:
                        continue
                    hyp = generated[base + src].clone()
                    score = s / (gen_len ** length_penalty)
                    if len(finished[i]) < nb or score > worst(i):
                        finished[i].append((score, torch.cat([hyp, torch.tensor([tok], device=dev)])))
                        if len(finished[i]) > nb:
                            finished[i].remove(min(finished[i], key=lambda h: h[0]))
                else:
                    keep.append((s, tok, base + src))
                if len(keep) == nb:
                    break
            for j, (s, tok, src) in enumerate(keep):
                next_scores[i, j], next_tokens[i, j], next_src[i, j] = s, tok, src
            if len(finished[i]) >= nb:
                if early_stopping is True:
                    done[i] = True
                else:
                    best_running = keep[0][0] if keep else -1e9
                    if early_stopping == "never" and length_penalty > 0:
                        bound = best_running / (max_new ** length_penalty)
                    else:
                        bound = best_running / (gen_len ** length_penalty)
                    done[i] = worst(i) >= bound
        beam_scores = next_scores.view(-1)
        src = next_src.view(-1)
        generated = torch.cat([generated[src], next_tokens.view(-1, 1)], dim=1)
        if all(done) or step + 1 == max_new:
            break
        stepper.reorder(src)
        logits = (graph if graph is not None else stepper).step(next_tokens.view(-1))

    out = []
    for i in range(batch):
        if not done[i]:  # add the running beams as finished hypotheses (HF finalize)
            for j in range(nb):
                s = float(beam_scores[i * nb + j])
                hyp = generated[i * nb + j]
                score = s / (hyp.shape[0] ** length_penalty)
                if len(finished[i]) < nb or score > worst(i):
                    finished[i].append((score, hyp))
                    if len(finished[i]) > nb:
                        finished[i].remove(min(finished[i], key=lambda h: h[0]))
        out.append(max(finished[i], key=lambda h: h[0])[1])
    width = max(h.shape[0] for h in out)
    res = torch.full((batch, width), pad_id, dtype=torch.long, device=dev)
    for i, h in enumerate(out):
        res[i, : h.shape[0]] = h
    return res
