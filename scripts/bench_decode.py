"""Decode benchmark (BASELINE.json configs[4]): greedy decode after a 16-context prompt
(17 clips x 8 frames, L = 958), full-size random-init model, batch sweep.  The decode loop
(CUDA-graphed step + device argmax) is timed with CUDA events over 64 steps; vision tower
and prefill are excluded.  Prints tok/s and the HBM-roofline fraction per batch size."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [1]
    cfg = bench.full_config()
    dev = torch.device("cuda", 0)
    model = bench.build_gpu_model(cfg, dev).eval()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    out = []
    for b in batches:
        for rep in range(2):
            d = bench.measure_decode(model, dev, batch=b)
        d["hbm_frac"] = d["bytes_per_token"] / (d["ms_per_token"] * 1e-3) / 1e9 / hbm
        out.append(d)
        print(json.dumps(d), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "bench_decode.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
