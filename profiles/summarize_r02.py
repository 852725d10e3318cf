"""Write profiles/r02_edge_kernels_ncu.json — the counters bench.py's roofline object reads (no literals in bench.py).

    python profiles/summarize_r02.py gpurun_out/r02_s2_prof.ncu-rep

Input: the `ncu --set full` capture of the two edge kernels of ONE pass of `bench.py --steps 2 --warmup 3 --no-cpu
--no-sdrf --no-dense --no-clocks` (arxiv-shaped graph).  Streamed entries come from the CPU model of the cheaper-side
rule (profiles/stream_model.py; no GPU needed).
"""
import csv
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "discrete-curvature-rewiring_b200"))
from dcr.synth import csr_from_edge_index, named_graph  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1.0}


def main(src):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(name):
        i = hdr.index(name)
        return [float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0) for r in data]

    names = [r[hdr.index("Kernel Name")] for r in data]
    dram = [a + b for a, b in zip(col("dram__bytes_read.sum"), col("dram__bytes_write.sum"))]
    inst = col("smsp__inst_executed.sum")
    dur = col("gpu__time_duration.sum")
    ei, n = named_graph("arxiv")
    rowptr, colidx = csr_from_edge_index(ei, n)
    deg = np.diff(rowptr.astype(np.int64))
    S = np.add.reduceat(deg[colidx], rowptr[:-1].astype(np.int64)) * (deg > 0)
    m = ei[0] < ei[1]
    I, J = ei[0][m], ei[1][m]
    stream = np.minimum(S[J] - deg[I], S[I] - deg[J])
    triv = np.minimum(deg[I], deg[J]) <= 1
    out = {
        "source": f"profiles/r02_ncu_full_edge_kernels.csv ({os.path.basename(src)}: ncu --set full --clock-control none, "
                  "one launch each of paper_group_kernel and paper_light_warp_kernel)",
        "kernels": names,
        "dram_bytes_per_kernel": dram,
        "dram_bytes_per_pass": float(sum(dram)),
        "warp_instructions_per_kernel": inst,
        "warp_instructions_per_pass": float(sum(inst)),
        "ncu_duration_per_kernel": dur,
        "streamed_entries_per_pass": int(stream[~triv].sum()),
        "streamed_entries_source": "CPU model of the cheaper-side rule (profiles/stream_model.py)",
    }
    with open(os.path.join(REPO, "profiles", "r02_edge_kernels_ncu.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
