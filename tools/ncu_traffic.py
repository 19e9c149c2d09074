"""profiles/rNN_dram_traffic.json from the raw page of the layer-set ncu capture
(tools/final_profiles.sh: `ncu --set full ... -k regex:fp4_gemm -s 12 -c 4 python bench.py --steps 1
--warmup 3 --no-details`, then `ncu -i ... --page raw --csv`).  The four captured launches are the
timed step's qkv, o, gate_up, down.

  python tools/ncu_traffic.py gpurun_out/final/ncu_full_layerset_raw.csv profiles/r02_dram_traffic.json
"""
import csv
import json
import sys

SHAPES = (("qkv", 10240, 8192), ("o", 8192, 8192), ("gate_up", 57344, 8192), ("down", 8192, 28672))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3, "%": 1.0,
         "register/thread": 1.0}


def main(src, dst, m=16):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(r, name):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)

    tensor = [h for h in hdr if h.startswith("sm__pipe_tensor") and "pct_of_peak_sustained_active" in h]
    out = {"source": "ncu --set full --clock-control none --import-source on -k regex:fp4_gemm -s 12 -c 4, "
                     "python bench.py --steps 1 --warmup 3 --no-details (raw page: "
                     + src.split("/")[-1] + "), final round-2 build", "per_launch": []}
    tot = 0.0
    for (name, n, k), r in zip(SHAPES, data):
        algo = n * k // 2 + n * k // 16 + 2 * m * k + 2 * m * n + 4
        rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
        tot += rd + wr
        e = {"gemm": name, "algorithmic_bytes": algo, "dram_read_bytes": rd, "dram_write_bytes": wr,
             "traffic_over_algorithmic": round((rd + wr) / algo, 4),
             "ncu_us": col(r, "gpu__time_duration.sum"),
             "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "alu_pipe_pct": col(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
             "fma_pipe_pct": col(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
             "registers_per_thread": col(r, "launch__registers_per_thread"),
             "kernel": r[hdr.index("Kernel Name")][:80]}
        if tensor:
            e["tensor_pipe_pct"] = col(r, tensor[0])
        out["per_launch"].append(e)
    out["avg_bytes_per_launch_m16"] = tot / len(out["per_launch"])
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps([(e["gemm"], e["traffic_over_algorithmic"], e["ncu_us"]) for e in out["per_launch"]]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
