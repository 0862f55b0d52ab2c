"""profiles/r02_ncu_gemm_traffic.json from the ncu report of scripts/profile_gemm_shapes.py:
    python scripts/gemm_traffic.py gpurun_out/r02_gemm_shapes.ncu-rep profiles/r02_ncu_gemm_traffic.json
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the i-th gemm_tcgen05 launch <-> SHAPES[i]."""
import csv
import json
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(rep, dst):
    shapes = [("vit.qkv", 34952, 4224, 1408), ("vit.proj", 34952, 1408, 1408), ("vit.fc1", 34952, 6144, 1408),
              ("vit.fc2", 34952, 1408, 6144), ("qf.crosskv", 34952, 9216, 1408)]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    gemms = [r for r in rows[2:] if "gemm_tcgen05" in r[col["Kernel Name"]]]
    for (name, m, n, k), r in zip(shapes, gemms):
        def val(key):
            return float(r[col[key]].replace(",", "")) * UNIT.get(units[col[key]], 1.0)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        launches.append(dict(name=name, mnk=[m, n, k], kernel=r[col["Kernel Name"]][:60], dram_bytes=rd + wr,
                             dram_read=rd, dram_write=wr, algorithmic_bytes=2.0 * (m * k + n * k + m * n),
                             duration_us=float(r[col["gpu__time_duration.sum"]].replace(",", ""))
                             * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[col["gpu__time_duration.sum"]], 1.0),
                             tensor_pipe_pct=float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])))
    Path(dst).write_text(json.dumps(dict(source=Path(rep).name, how="ncu --set full --clock-control none, cold L2, one "
                                         "launch per shape (scripts/profile_gemm_shapes.py)", launches=launches), indent=1))
    for l in launches:
        print(l)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
