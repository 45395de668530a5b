"""Numerical cost of the tensor-core operand schemes considered for the fp32-grade GEMMs, emulated on the
CPU: operands are rounded to the storage format plane by plane (round to nearest even), products are accumulated in
float64 (the TMEM accumulator's fp32 rounding is common to all schemes and far below these figures).  Shapes and value
scales follow the decoder's per-timestep GEMMs (K = 2048, weights U(+-1/sqrt(K)), activations in [-1, 1]) and a 3x3
convolution (K = 4608, post-ReLU activations with a large mean).  Error = max |C - C_ref| / max |C_ref| over the output.

    python tools/operand_schemes.py            # prints the table kept as profiles/r02_operand_schemes.txt
"""
import numpy as np


def rnd_bits(x, mant):
    """round float32 to `mant` explicit mantissa bits (bf16: 7, tf32: 10), nearest even, exponent range kept"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - mant
    bias = ((u >> drop) & 1) + ((1 << (drop - 1)) - 1)
    u = ((u + bias) >> drop) << drop
    return u.astype(np.uint32).view(np.float32)


FORMATS = {"bf16": lambda x: rnd_bits(x, 7), "fp16": lambda x: np.asarray(x, np.float32).astype(np.float16).astype(np.float32),
           "tf32": lambda x: rnd_bits(x, 10)}


def planes(x, fmt, n):
    out, r = [], np.asarray(x, np.float32)
    for _ in range(n):
        p = FORMATS[fmt](r)
        out.append(p.astype(np.float64))
        r = (r.astype(np.float64) - p).astype(np.float32)
    return out


def gemm(a_planes, b_planes, terms):
    return sum(a_planes[i] @ b_planes[j].T for i, j in terms)


SCHEMES = [  # name, format, planes (A, B), product terms (A plane, B plane), bytes per element (A, B), MMA issues per k-step
    ("bf16 (1 plane)", "bf16", (1, 1), [(0, 0)], (2, 2), 1.0),
    ("fp16 (1 plane)", "fp16", (1, 1), [(0, 0)], (2, 2), 1.0),
    ("tf32 (1 plane)", "tf32", (1, 1), [(0, 0)], (4, 4), 2.0),            # kind::tf32 runs at half the bf16 rate
    ("fp16 W(hi,lo) x fp16 x: 2 MMAs", "fp16", (2, 1), [(0, 0), (1, 0)], (4, 2), 2.0),
    ("bf16 W(hi,lo) x bf16 x: 2 MMAs", "bf16", (2, 1), [(0, 0), (1, 0)], (4, 2), 2.0),
    ("bf16x3: hi*hi + hi*lo + lo*hi (shipped)", "bf16", (2, 2), [(0, 0), (0, 1), (1, 0)], (4, 4), 3.0),
    ("fp16x3: hi*hi + hi*lo + lo*hi", "fp16", (2, 2), [(0, 0), (0, 1), (1, 0)], (4, 4), 3.0),
    ("tf32x3", "tf32", (2, 2), [(0, 0), (0, 1), (1, 0)], (8, 8), 6.0),
    ("bf16x4: all four products", "bf16", (2, 2), [(0, 0), (0, 1), (1, 0), (1, 1)], (4, 4), 4.0),
]


def cases(rng):
    K = 2048
    yield ("decoder layer GEMM: 4096 x 64, K = 2048 (tanh-range activations)",
           rng.uniform(-1, 1, (4096, K)).astype(np.float32) / np.sqrt(K), np.tanh(rng.standard_normal((64, K))).astype(np.float32))
    K = 4608
    yield ("conv6 as a GEMM: 512 x 256 pixels, K = 4608 (post-ReLU activations, mean >> 0)",
           rng.uniform(-1, 1, (512, K)).astype(np.float32) / np.sqrt(K),
           np.maximum(rng.standard_normal((256, K)) * 0.5 + 1.0, 0).astype(np.float32))


def main():
    rng = np.random.default_rng(910820)
    for title, A, B in cases(rng):
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        print(title)
        print(f"  {'scheme':44s} {'max rel err':>12s} {'B/elem A,B':>11s} {'MMA issues':>10s}")
        for name, fmt, (na, nb), terms, bytes_, mmas in SCHEMES:
            C = gemm(planes(A, fmt, na), planes(B, fmt, nb), terms)
            err = np.abs(C - ref).max() / np.abs(ref).max()
            print(f"  {name:44s} {err:12.2e} {bytes_[0]:>5d},{bytes_[1]:<5d} {mmas:10.1f}")
        print()
    print("Bar: logits within 1e-3 and greedy tokens exact outside ties after 7 convolution layers, 24 encoder and 20-50\n"
          "decoder steps of dependent GEMMs; random-init top-1 / top-2 log-prob gaps are ~1e-3.  Single-plane schemes and the\n"
          "2-MMA schemes (single-plane activations) sit at 1e-4 .. 3e-3 PER GEMM; only the two-plane x two-plane schemes reach\n"
          "1e-5 or better, and among those bf16x3 has the smallest footprint at the full bf16 MMA rate (fp16 pairs: same bytes,\n"
          "same MMA count, 3x better error but a 5-bit exponent that the 1/B-scaled gradients (1e-6 .. 1e-9) underflow or\n"
          "flush; tf32 pairs: twice the bytes, half the rate).")


if __name__ == "__main__":
    main()
