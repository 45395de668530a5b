"""Turns the raw ncu CSV pages of tools/gpu_job_ncu.sh (gpurun_out/<tag>_executor_raw.csv, <tag>_gemm_metrics.csv) into
the summaries kept under profiles/ (<tag>_executor_ncu_full.json, <tag>_tc_gemm_ncu_full.json)."""
import collections
import csv
import json
import re
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
SCALE_B = {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}
SCALE_T = {"us": 1.0, "ns": 1e-3, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def num(v):
    v = v.replace(",", "")
    return float(v) if v not in ("", "n/a") else None


rows = list(csv.reader(open(f"gpurun_out/{tag}_executor_raw.csv")))
hdr, units = rows[0], rows[1]


def col(name):
    if name in hdr:
        return hdr.index(name)
    return [i for i, h in enumerate(hdr) if h.endswith(name)][0]


names = {"65": "enc_bwd, one direction (GEMM + cell-backward commands per step)",
         "1057": "enc_fwd, one direction (one fused GEMM->cell command per step)",
         "2059": "dec_fwd (per step: 2 layer GEMM+cell, stacked [W_a;W_c2] GEMM, attention+output)",
         "149": "dec_bwd (per step: attention backward + 3 GEMM + 2 cell-backward commands)",
         "6923": "dual greedy decode (greedy + gold rows in one batch)"}
want = {"us": "gpu__time_duration.sum", "dram_read_mb": "dram__bytes_read.sum", "dram_write_mb": "dram__bytes_write.sum",
        "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pipe_active_pct": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "registers": "launch__registers_per_thread"}
ks = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    m = re.search(r"persist_kernel<(\d+), (\d+)>", r[col("Kernel Name")])
    d = {"kernel": f"persist_kernel<{m.group(1)}, {m.group(2)}>", "program": names.get(m.group(2), "?"), "grid": r[col("Grid Size")]}
    for k, c in want.items():
        x, u = num(r[col(c)]), units[col(c)]
        if x is not None and k == "us":
            x *= SCALE_T[u]
        if x is not None and k.endswith("_mb"):
            x *= SCALE_B[u]
        d[k] = None if x is None else round(x, 4)
    ks.append(d)
# one whole train step = enc_fwd x2, dec_fwd, dec_bwd, enc_bwd x2 in launch order
sets = [k["kernel"].split(", ")[1].rstrip(">") for k in ks]
start = next(i for i in range(len(ks) - 5) if sets[i:i + 6] == ["1057", "1057", "2059", "149", "65", "65"])
step = ks[start:start + 6]
tot = sum(k["dram_read_mb"] + k["dram_write_mb"] for k in step)
json.dump({"command": "AOCR_GRAPHS=0 STEP_N=3 STEP_DECODE=1 ncu --set full --clock-control none -k regex:persist_kernel -s 12 -c 9 "
                      "python tools/one_step.py  (tools/gpu_job_ncu.sh)",
           "note": "The executor as shipped: under Nsight Compute the cooperative + thread-block-cluster launch is rejected; the "
                   "library's launch-mode probe then launches the same kernel cooperatively without clusters (the GEMM -> cell "
                   "pairs run as two commands). ncu serialises kernels and replays them with cold caches: durations here are not "
                   "bench numbers.",
           "what": f"executor launches around one config-2 train step (batch 64, S=24, T=20); entries {start}..{start + 5} are one whole step",
           "dram_bytes_per_launch": round(tot * 1e6 / 6), "dram_mb_per_train_step": round(tot, 1), "kernels": ks},
          open(f"profiles/{tag}_executor_ncu_full.json", "w"), indent=1)
print("executor: DRAM MB per train step", round(tot, 1), "-> per launch", round(tot / 6, 1))

lines = [l for l in open(f"gpurun_out/{tag}_gemm_metrics.csv") if l.startswith('"')]
by = collections.OrderedDict()
for r in csv.DictReader(lines):
    km = re.search(r"(tc_gemm2?p?_kernel)<(?:\(int\))?(\d+)>", r["Kernel Name"])
    e = by.setdefault(r["ID"], {"kernel": f"{km.group(1)}<{km.group(2)}>" if km else r["Kernel Name"][:60], "grid": r.get("Grid Size")})
    e[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
gl = []
for d in by.values():
    def val(n):
        return num(d[n][0])

    def mb(n):
        return num(d[n][0]) * SCALE_B[d[n][1]]
    us = num(d["gpu__time_duration.sum"][0]) * SCALE_T[d["gpu__time_duration.sum"][1]]
    gl.append({"kernel": d["kernel"], "grid": d["grid"], "us": round(us, 2), "dram_read_mb": round(mb("dram__bytes_read.sum"), 3),
               "dram_write_mb": round(mb("dram__bytes_write.sum"), 3),
               "tensor_pipe_active_pct": val("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
               "l2_throughput_pct": val("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "sm_throughput_pct": val("sm__throughput.avg.pct_of_peak_sustained_elapsed"), "registers": int(val("launch__registers_per_thread"))})
tt = sum(g["us"] for g in gl)
wt = sum(g["us"] * (g["tensor_pipe_active_pct"] or 0.0) for g in gl) / tt
json.dump({"command": "AOCR_GRAPHS=0 STEP_N=2 ncu --metrics <duration, dram bytes, tensor pipe, L2, SM throughput> --clock-control none "
                      "-k regex:tc_gemm -s 40 -c 40 python tools/one_step.py  (tools/gpu_job_ncu.sh)",
           "what": "the tcgen05 GEMM / implicit-GEMM convolution launches of one config-2 train step in launch order (forward "
                   "convolutions, encoder input projection, ..., weight and data gradients); tc_gemm2_kernel = CTA-pair kernel",
           "time_weighted_tensor_pipe_active_pct": round(wt, 2), "total_us": round(tt, 1),
           "dram_bytes_per_launch": round(sum(g["dram_read_mb"] + g["dram_write_mb"] for g in gl) * 1e6 / len(gl)), "launches": gl},
          open(f"profiles/{tag}_tc_gemm_ncu_full.json", "w"), indent=1)
print("gemm: launches", len(gl), "total us", round(tt, 1), "time-weighted tensor pipe active %", round(wt, 2))
for g in gl:
    print(f"  {g['kernel'][:22]:22s} grid {g['grid']:>14s} {g['us']:8.2f} us  tensor {g['tensor_pipe_active_pct']}  L2 {g['l2_throughput_pct']}  dram rd {g['dram_read_mb']} MB")
