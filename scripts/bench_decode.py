"""Decode benchmark (BASELINE.json configs[4]): greedy generate() of 32 new tokens after a
16-context prompt (17 clips x 8 frames, L = 958), batch sweep, full-size random-init model.
Reports tok/s of the decode loop (prefill excluded) and the HBM roofline fraction."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from eilev_b200.engine import opt as E_opt  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [1]
    cfg = bench.full_config()
    dev = torch.device("cuda", 0)
    model = bench.build_gpu_model(cfg, dev).eval()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    out = []
    for b in batches:
        one = bench.synthetic_batch(7)
        n = int(one["attention_mask"].sum()) - bench.TARGET_TOKENS  # prompt without the target
        ids = one["input_ids"][:, :n].repeat(b, 1).to(dev)
        vm = one["video_input_mask"][:, :n].repeat(b, 1).to(dev)
        am = torch.ones_like(ids)
        px = one["pixel_values"].to(dev)
        px = px.repeat(b, 1, 1, 1, 1) if b > 1 else px
        def run(new, graph):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            toks = model.generate(ids, pixel_values=px, video_input_mask=vm, attention_mask=am,
                                  max_new_tokens=new, min_new_tokens=new, do_sample=False, eos_token_id=None,
                                  use_cuda_graph=graph)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, toks

        run(2, False)  # warm-up (weight packing)
        run(12, True)  # warm-up of the graph path (first instantiate is slow)
        for graph in (True, False):
            # per-token time = slope between two generation lengths: fixed costs (vision tower,
            # prefill, graph capture) cancel
            t_short, _ = run(16, graph)
            t_long, toks = run(80, graph)
            per_tok = (t_long - t_short) / 64
            tok_s = b / per_tok
            bytes_step = 5.293e9 + 327680.0 * (n + 48) * b
            row = dict(batch=b, prompt_len=n, cuda_graph=graph, fixed_ms=round((t_short - 16 * per_tok) * 1e3, 1),
                       decode_ms_per_token=round(per_tok * 1e3, 3), tok_s=round(tok_s, 1),
                       hbm_gbs=round(bytes_step / per_tok / 1e9, 1),
                       hbm_frac=round(bytes_step / per_tok / (hbm * 1e9), 3), shape=list(toks.shape))
            out.append(row)
            print(json.dumps(row), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "bench_decode.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
