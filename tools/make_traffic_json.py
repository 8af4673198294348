"""profiles/r02_traffic.json (what bench.py cites as `roofline.traffic`) and profiles/r02_step_ncu_full_summary.csv from
one `ncu --set full` capture of a headline step, read here with `ncu -i <rep> --page raw --csv`:

    ncu --set full --clock-control none --import-source on \
        -k regex:"digit_gemm_pair_persistent|feature_max_tiled|digit_split|counts_window|moments_pass|digit_combine" \
        -c 14 -o gpurun_out/r02_step_full \
        python bench.py --steps 2 --warmup 1 --no-fp64-leg --no-sample --no-other-configs --no-cpu-baseline
    python tools/make_traffic_json.py gpurun_out/r02_step_full.ncu-rep

The traffic of the integer engine's GEMM is the SUM over the launches of one step (one launch per batch of resamples);
sample and FP64 entries are kept from the captures named in their `source`.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(rep, batches_per_step=4):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    out = os.path.join(ROOT, "profiles", "r02_step_ncu_full_summary.csv")
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none of the kernels of one headline step (command in "
                "tools/make_traffic_json.py); D=64 N=2^20 R=10^4 second order, 4 GEMM batches of 2560 resamples, batches 0-2 host the "
                "next batch's Philox draw; cold-cache, serialised launches\n")
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i][:110] if hdr[i] == "Kernel Name" else r[i] for i in idx])
    name, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

    def gb(r, i):
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[units[i]]
        return float(r[i].replace(",", "")) * scale

    gemm = [r for r in data if "digit_gemm_pair_persistent" in r[name]][:batches_per_step]
    assert len(gemm) == batches_per_step, f"expected {batches_per_step} GEMM launches of one step, found {len(gemm)}"
    total = sum(gb(r, rd) + gb(r, wr) for r in gemm)
    per = ", ".join(f"{(gb(r, rd) + gb(r, wr)) / 1e9:.1f}" for r in gemm)
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    with open(path) as f:
        t = json.load(f)
    t["digits_d64_n20_r10000_c1_g1"] = {
        "bytes": int(total),
        "source": "profiles/r02_step_ncu_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                  f"{batches_per_step} launches of digit_gemm_pair_persistent_kernel in one step ({per} GB; a launch "
                  "that hosts a draw also writes the next batch's 2.7 GB of multiplicities); operands of one launch: "
                  "W batch 2.7 + Q 15.8 + C 0.15 = 18.6 GB, i.e. 74.5 GB per step as batched, 26.9 GB if Q were read once",
        "launches_per_step": batches_per_step}
    with open(path, "w") as f:
        json.dump(t, f, indent=1)
    print("wrote", out, "and", path, f"(GEMM traffic per step {total / 1e9:.1f} GB)")


if __name__ == "__main__":
    main(sys.argv[1])
