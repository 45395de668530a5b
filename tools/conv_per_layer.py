"""Per-layer table of the convolution GEMMs of one config-2 train step from profiles/<tag>_tc_gemm_ncu_full.json
(tools/gpu_job_ncu.sh + tools/ncu_to_profiles.py) -> profiles/<tag>_conv_per_layer.md.  The launches of the capture window are
matched to (layer, role) by their position in the step's launch order (forward convolutions conv2..7, then the backward
chain conv7..2 with the weight gradient of a layer launched before its data gradient)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
d = json.load(open(f"profiles/{tag}_tc_gemm_ncu_full.json"))
L = d["launches"]
B = 64
conv = {2: (64, 128, 3, 16, 50), 3: (128, 256, 3, 8, 25), 4: (256, 256, 3, 8, 25), 5: (256, 512, 3, 4, 25), 6: (512, 512, 3, 4, 25),
        7: (512, 512, 2, 1, 24)}
hin = {2: (16, 50), 3: (8, 25), 4: (8, 25), 5: (4, 25), 6: (4, 25), 7: (2, 25)}
idx = {("fwd", 2): 32, ("fwd", 3): 33, ("fwd", 4): 34, ("fwd", 5): 35, ("fwd", 6): 36, ("fwd", 7): 37,
       ("wgrad", 6): 22, ("dgrad", 6): 23, ("wgrad", 5): 24, ("dgrad", 5): 25, ("wgrad", 4): 26, ("dgrad", 4): 27,
       ("wgrad", 3): 28, ("dgrad", 3): 29, ("wgrad", 2): 30, ("dgrad", 2): 31, ("dgrad/wgrad", 7): 21}
rows = []
for (role, l), i in sorted(idx.items(), key=lambda kv: (kv[0][1], kv[0][0])):
    cin, cout, k, ho, wo = conv[l]
    hi, wi = hin[l]
    pix = B * ho * wo
    flop = 2.0 * pix * cout * k * k * cin
    act_in, w = B * hi * wi * cin * 4, cout * k * k * cin * 4     # bf16 plane pairs: 4 B per element
    algb = act_in + w + pix * cout * 4                            # two operands read once + one fp32 / plane-pair result written
    g = L[i]
    rows.append((f"conv{l} {role}", g["kernel"], g["grid"], g["us"], flop / 1e9, flop / g["us"] / 1e6, 3 * flop / g["us"] / 1e6,
                 g["tensor_pipe_active_pct"], g["dram_read_mb"] + g["dram_write_mb"], algb / 1e6))
out = ["# Convolution GEMMs of one config-2 train step (batch 64), per layer\n",
       f"Source: `{tag}_tc_gemm_ncu_full.json` (`ncu` per-launch metrics, kernels serialised with cold caches: a launch here runs alone,",
       "in the live step the weight-gradient lane overlaps the data-gradient chain).  FLOP = 2 x pixels x Cout x k x k x Cin (each MAC once);",
       "the bf16x3 mode issues 3 MMAs per MAC, so `MMA TFLOP/s` = 3 x algorithmic.  Algorithmic bytes = both operands as bf16 plane pairs",
       "(4 B per element) read once + the result written once; DRAM = `dram__bytes_read + dram__bytes_write` of the launch",
       "(less than the algorithmic figure where an operand was still in L2 from its producer).  `tc_gemm2p_kernel` = persistent pair kernel",
       "(grid = 148 CTAs = 74 pairs looping over the tiles).\n",
       "| layer, role | kernel | grid | us | GFLOP | TFLOP/s | MMA TFLOP/s | tensor pipe active % | DRAM MB | algorithmic MB |",
       "|---|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    out.append(f"| {r[0]} | `{r[1]}` | {r[2]} | {r[3]:.1f} | {r[4]:.2f} | {r[5]:.0f} | {r[6]:.0f} | {r[7]:.1f} | {r[8]:.1f} | {r[9]:.1f} |")
out.append("\nconv7 (2 x 2, no padding, 24 output columns): one of its two backward GEMMs fell inside the capture window.  conv2's weight")
out.append("gradient has M = Cout = 128 (one M tile) and runs on the single-CTA kernel with split-K 16.")
open(f"profiles/{tag}_conv_per_layer.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[9:]))
