"""What an optimizer step costs next to the micro-steps around it (opt-2.7b, full size, CUDA graphs): CUDA-event
times of 34 consecutive micro-steps; every 16th carries the optimizer step, the one after it replays the graph that
re-packs the Q-Former."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from eilev_b200.train import DataParallelTrainer  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = bench.full_config(0.1, "opt")
model = bench.build_gpu_model(cfg, dev)
trainer = DataParallelTrainer(model, lr=1e-5, weight_decay=0.05, max_grad_norm=1.0, grad_accum=bench.GRAD_ACCUM)
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(1000).items()}
trainer.capture_graph(batch)
for _ in range(3):
    trainer.micro_step(batch)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(35)]
ev[0].record()
marks = []
for i in range(34):
    trainer.micro_step(batch)
    marks.append(trainer.micro % trainer.grad_accum)
    ev[i + 1].record()
torch.cuda.synchronize()
for i in range(34):
    tag = "  <- optimizer step at the end" if marks[i] == 0 else ("  <- re-pack graph" if marks[i] == 1 else "")
    print(f"micro-step {i:2d} (accum {marks[i]:2d}): {ev[i].elapsed_time(ev[i + 1]):7.2f} ms{tag}")
print(f"mean of all 34: {ev[0].elapsed_time(ev[34]) / 34:.2f} ms")
