#!/usr/bin/env python3
"""Emit fft64x2_gen.cuh: the in-register 64-point complex forward DFT of gen_fft64.py, written with Blackwell's packed
FP32 pair instructions (PTX add/sub/mul.rn.f32x2 -> SASS FADD2 / FMUL2, sm_100+).

    python audio-training_b200/csrc/gen_fft64x2.py > audio-training_b200/csrc/fft64x2_gen.cuh

Why: measured on B200 (tools/ubench_f32x2.cu) FADD2 issues at half the rate of FADD -- the same flops per clock --
but takes one issue slot for two lanes, reads its operands as aligned even/odd register pairs (no register-bank
conflicts between the two halves) and halves the code size.  The scalar kernel spent 15-20 % of its FFT cycles in
dispatch stalls and 20 % waiting for instructions.

Same 8 x 8 Cooley-Tukey as the scalar generator (n = 8a + b, k = k1 + 8 k2):
  pass 1: for each b a radix-8 DFT over a.  Butterflies b = 2 beta and 2 beta + 1 ride the two halves of one pair:
          the pair (x[8a + 2 beta], x[8a + 2 beta + 1]) is two ADJACENT inputs, which is how the callers produce them
          (a 16-byte shared-memory load yields two pairs; the window multiply is itself a packed multiply);
  twiddle: slot (k1, b) *= W64^(b k1), scalar (literal constants differ per half) -- its results are written straight
          into the pairing of pass 2, so no register shuffling is needed for the 49 non-trivial elements;
  pass 2: for each k1 a radix-8 DFT over b; butterflies k1 = 2 kappa and 2 kappa + 1 share a pair, so the outputs
          X[2 kappa + 8 k2], X[2 kappa + 1 + 8 k2] are adjacent frequencies.
X[k] ends up in re[k]/im[k] (natural order, unlike cacfe_fft64): inputs (n, n+1) and outputs (k, k+1) share register
pairs, so a loop around this function needs no register shuffling.
"""
import math

SQ = "0.70710678118654752440f"


class Gen:
    def __init__(self):
        self.lines = []
        self.n = 0

    def new(self, prefix="p"):
        self.n += 1
        return f"{prefix}{self.n}"

    def emit(self, s):
        self.lines.append("  " + s)

    def add(self, a, b):
        d = self.new()
        self.emit(f"const cacfe_f2 {d} = cacfe_add2({a}, {b});")
        return d

    def sub(self, a, b):
        d = self.new()
        self.emit(f"const cacfe_f2 {d} = cacfe_sub2({a}, {b});")
        return d

    def mulsq(self, a):
        d = self.new()
        self.emit(f"const cacfe_f2 {d} = cacfe_mul2({a}, sq2);")
        return d

    def fmasq(self, a, c, neg=False):
        """c + a / sqrt2 (or c - a / sqrt2): one packed FMA with a constant operand -- it costs what the packed add costs
        (tools/ubench_issue.cu: 2.07 cycles against 2.08) and the multiply by 1 / sqrt2 disappears."""
        d = self.new()
        self.emit(f"const cacfe_f2 {d} = cacfe_fma2({a}, {'nsq2' if neg else 'sq2'}, {c});")
        return d


def radix8(g, xr, xi):
    """Packed natural-order radix-8 DIF butterfly.  xr/xi: 8 packed names each.  Returns (yr, yi) lists for k = 0..7.
    Negations are tracked symbolically (sign, name) so that multiplications by -i cost nothing."""
    ar, ai, br, bi = [], [], [], []
    for i in range(4):
        ar.append(g.add(xr[i], xr[i + 4]))
        ai.append(g.add(xi[i], xi[i + 4]))
        br.append(g.sub(xr[i], xr[i + 4]))
        bi.append(g.sub(xi[i], xi[i + 4]))
    # odd branch: c0 = b0; c1 = b1 * (1 - i)/sqrt2; c2 = b2 * (-i); c3 = b3 * (-1 - i)/sqrt2.  The factor 1 / sqrt2 of c1 and c3 is
    # folded into the last stage (FOLD_SQ2): u1 = sqrt2 c1, u3 = sqrt2 c3 (with the sign convention n3i of the plain form)
    u1r = g.add(br[1], bi[1])
    u1i = g.sub(bi[1], br[1])
    u3r = g.sub(bi[3], br[3])
    m3i = g.add(br[3], bi[3])

    def dft4(s0r, s0i, s1r, s1i, s2r, s2i, dr, di):
        """outputs: y0 = s0 + s2, y2 = s0 - s2, y1 = s1 + (-i) d = (s1r + di, s1i - dr), y3 = (s1r - di, s1i + dr)"""
        return [(g.add(s0r, s2r), g.add(s0i, s2i)), (g.add(s1r, di), g.sub(s1i, dr)),
                (g.sub(s0r, s2r), g.sub(s0i, s2i)), (g.sub(s1r, di), g.add(s1i, dr))]

    def dft4_sq(s0r, s0i, s1r, s1i, s2r, s2i, dr, di):
        """the same with s2 and d still carrying a factor sqrt2: eight packed FMAs with the constant +- 1 / sqrt2"""
        return [(g.fmasq(s2r, s0r), g.fmasq(s2i, s0i)), (g.fmasq(di, s1r), g.fmasq(dr, s1i, True)),
                (g.fmasq(s2r, s0r, True), g.fmasq(s2i, s0i, True)), (g.fmasq(di, s1r, True), g.fmasq(dr, s1i))]

    # even branch on a0..a3
    e = dft4(g.add(ar[0], ar[2]), g.add(ai[0], ai[2]), g.sub(ar[0], ar[2]), g.sub(ai[0], ai[2]),
             g.add(ar[1], ar[3]), g.add(ai[1], ai[3]), g.sub(ar[1], ar[3]), g.sub(ai[1], ai[3]))
    # odd branch on c0..c3 with c2 = (b2i, -b2r), sqrt2 c1 = (u1r, u1i), sqrt2 c3 = (u3r, -m3i)
    s0r, s0i = g.add(br[0], bi[2]), g.sub(bi[0], br[2])      # c0 + c2
    s1r, s1i = g.sub(br[0], bi[2]), g.add(bi[0], br[2])      # c0 - c2
    s2r, s2i = g.add(u1r, u3r), g.sub(u1i, m3i)              # sqrt2 (c1 + c3)
    dr, di = g.sub(u1r, u3r), g.add(u1i, m3i)                # sqrt2 (c1 - c3)
    o = dft4_sq(s0r, s0i, s1r, s1i, s2r, s2i, dr, di)
    yr = [None] * 8
    yi = [None] * 8
    for q in range(4):
        yr[2 * q], yi[2 * q] = e[q]
        yr[2 * q + 1], yi[2 * q + 1] = o[q]
    return yr, yi


def twiddle_scalar(g, r, i, m):
    """(r + i i) * W64^m on scalars; returns new (r, i) names."""
    m %= 64
    if m == 0:
        return r, i
    nr, ni = g.new("t"), g.new("t")
    if m % 16 == 0:
        q = m // 16
        if q == 1:
            g.emit(f"const float {nr} = {i}, {ni} = -{r};")
        elif q == 2:
            g.emit(f"const float {nr} = -{r}, {ni} = -{i};")
        else:
            g.emit(f"const float {nr} = -{i}, {ni} = {r};")
        return nr, ni
    c = math.cos(-2 * math.pi * m / 64)
    s = math.sin(-2 * math.pi * m / 64)
    if m % 8 == 0:
        sc = "" if c > 0 else "-"
        ss = "" if s > 0 else "-"
        g.emit(f"const float {nr} = ({sc}{r} - ({ss}{i})) * {SQ}, {ni} = ({ss}{r} + ({sc}{i})) * {SQ};")
        return nr, ni
    g.emit(f"const float {nr} = fmaf(-{i}, {s!r}f, {r} * {c!r}f), {ni} = fmaf({r}, {s!r}f, {i} * {c!r}f);")
    return nr, ni


def main():
    g = Gen()
    out = []
    out.append("// GENERATED by gen_fft64x2.py -- do not edit.  In-register 64-point forward DFT (8x8), packed f32x2 form.")
    out.append("#pragma once")
    out.append('#include "fft64_gen.cuh"   // CACFE_FFT64_SLOT, CACFE_HD')
    out.append("// Packed pair of floats.  Device: one 64-bit register (even/odd pair) driven by add/sub/mul.rn.f32x2;")
    out.append("// host (tests/emul): two floats with the same IEEE operations.")
    out.append("#ifdef __CUDA_ARCH__")
    out.append("typedef unsigned long long cacfe_f2;")
    out.append('__device__ __forceinline__ cacfe_f2 cacfe_pk(float lo, float hi) { cacfe_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }')
    out.append('__device__ __forceinline__ float cacfe_lo(cacfe_f2 v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }')
    out.append('__device__ __forceinline__ float cacfe_hi(cacfe_f2 v) { return __uint_as_float((unsigned)(v >> 32)); }')
    out.append('__device__ __forceinline__ cacfe_f2 cacfe_add2(cacfe_f2 a, cacfe_f2 b) { cacfe_f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }')
    out.append('__device__ __forceinline__ cacfe_f2 cacfe_sub2(cacfe_f2 a, cacfe_f2 b) { cacfe_f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }')
    out.append('__device__ __forceinline__ cacfe_f2 cacfe_mul2(cacfe_f2 a, cacfe_f2 b) { cacfe_f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }')
    out.append('__device__ __forceinline__ cacfe_f2 cacfe_fma2(cacfe_f2 a, cacfe_f2 b, cacfe_f2 c) { cacfe_f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }')
    out.append("#else")
    out.append("#include <cmath>")
    out.append("struct cacfe_f2 { float lo, hi; };")
    out.append("inline cacfe_f2 cacfe_pk(float lo, float hi) { return cacfe_f2{lo, hi}; }")
    out.append("inline float cacfe_lo(cacfe_f2 v) { return v.lo; }")
    out.append("inline float cacfe_hi(cacfe_f2 v) { return v.hi; }")
    out.append("inline cacfe_f2 cacfe_add2(cacfe_f2 a, cacfe_f2 b) { return cacfe_f2{a.lo + b.lo, a.hi + b.hi}; }")
    out.append("inline cacfe_f2 cacfe_sub2(cacfe_f2 a, cacfe_f2 b) { return cacfe_f2{a.lo - b.lo, a.hi - b.hi}; }")
    out.append("inline cacfe_f2 cacfe_mul2(cacfe_f2 a, cacfe_f2 b) { return cacfe_f2{a.lo * b.lo, a.hi * b.hi}; }")
    out.append("inline cacfe_f2 cacfe_fma2(cacfe_f2 a, cacfe_f2 b, cacfe_f2 c) { return cacfe_f2{std::fmaf(a.lo, b.lo, c.lo), std::fmaf(a.hi, b.hi, c.hi)}; }")
    out.append("#endif")
    out.append("CACFE_HD void cacfe_fft64x2(float (&re)[64], float (&im)[64]) {")
    g.emit(f"const cacfe_f2 sq2 = cacfe_pk({SQ}, {SQ}), nsq2 = cacfe_pk(-{SQ}, -{SQ});")
    # pass 1
    sr = {}  # (k1, b) -> scalar names after pass 1
    si = {}
    for beta in range(4):
        b = 2 * beta
        g.emit(f"// pass 1, b = {b}, {b + 1}")
        xr, xi = [], []
        for a in range(8):
            nr, ni = g.new("x"), g.new("x")
            g.emit(f"const cacfe_f2 {nr} = cacfe_pk(re[{8 * a + b}], re[{8 * a + b + 1}]), {ni} = cacfe_pk(im[{8 * a + b}], im[{8 * a + b + 1}]);")
            xr.append(nr)
            xi.append(ni)
        yr, yi = radix8(g, xr, xi)
        for k1 in range(8):
            for h, bb in ((0, b), (1, b + 1)):
                f = "cacfe_lo" if h == 0 else "cacfe_hi"
                nr, ni = g.new("s"), g.new("s")
                g.emit(f"const float {nr} = {f}({yr[k1]}), {ni} = {f}({yi[k1]});")
                sr[(k1, bb)], si[(k1, bb)] = nr, ni
    # twiddle (scalar), then pass 2
    g.emit("// twiddles W64^(b k1)")
    for k1 in range(8):
        for b in range(8):
            sr[(k1, b)], si[(k1, b)] = twiddle_scalar(g, sr[(k1, b)], si[(k1, b)], b * k1)
    for kappa in range(4):
        k1 = 2 * kappa
        g.emit(f"// pass 2, k1 = {k1}, {k1 + 1}")
        xr, xi = [], []
        for b in range(8):
            nr, ni = g.new("x"), g.new("x")
            g.emit(f"const cacfe_f2 {nr} = cacfe_pk({sr[(k1, b)]}, {sr[(k1 + 1, b)]}), {ni} = cacfe_pk({si[(k1, b)]}, {si[(k1 + 1, b)]});")
            xr.append(nr)
            xi.append(ni)
        yr, yi = radix8(g, xr, xi)
        for k2 in range(8):
            k = k1 + 8 * k2  # natural order: the pair (X[k], X[k+1]) stays in the register pair of inputs (k, k+1)
            g.emit(f"re[{k}] = cacfe_lo({yr[k2]}); im[{k}] = cacfe_lo({yi[k2]}); "
                   f"re[{k + 1}] = cacfe_hi({yr[k2]}); im[{k + 1}] = cacfe_hi({yi[k2]});")
    out += g.lines
    out.append("}")
    print("\n".join(out))


if __name__ == "__main__":
    main()
