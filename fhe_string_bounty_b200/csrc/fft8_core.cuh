// fft8_core.cuh -- the size-1024 negacyclic-carrier FFT once more, for 128 threads x 8 points: four warps per polynomial.
// Used by the latency-oriented blind-rotation instance (pbs_v8.cu): a narrow tree level puts ONE ciphertext on an SM, and with
// 16 points per thread (pbs_v4.cu) that is one warp per scheduler -- ncu: issue slots 26 % busy, the warp waits on its own
// instruction fetch and dependency latencies.  Half the work per warp and two warps per scheduler overlap those.
//
// Point j = T + 128*m (thread T = 0..127, register m = 0..7).  Z_k = sum_j z_j w^j W^(jk), k = k1 + 8*k2, k2 = s + 8*(e + 4*f):
//   pass 1   in-register radix-8 DIF over m of z * w^(128 m)                        -> register p1 = brev3(k1)
//   twiddle  T1[p1][T] = w^T * W^(T * brev3(p1))
//   exchange A: (T = a + 16b, p1) -> thread T' = a + 16*p1, register b             (128-point DFT over T remains)
//   pass 2   radix-8 DIF over b                                                     -> register ps = brev3(s)
//   twiddle  W128^(a * s): seven per-thread constants (a = T' & 15)
//   exchange B: (T' = a + 16*p1, ps), a = cl + 2c' + 4d -> thread T'' = c' + 2*ps + 16*p1, register d + 4*cl
//   pass 3   two radix-4 DIFs over d                                                -> register pe + 4*cl, e = brev2(pe)
//   twiddle  W16^((cl + 2c') * e): per-thread constants
//   exchange C: between the two threads c' = 0, 1 of a pair: (c', cl, pe = pe0 + 2*pe1) -> thread e' = pe1, register c + 4*pe0, c = cl + 2c'
//   pass 4   two radix-4 DIFs over c                                                -> register pf + 4*pe0, f = brev2(pf)
// Thread T'' = e' + 2*ps + 16*p1, register pf + 4*pe0 holds k = brev3(p1) + 8*(brev3(ps) + 8*(brev2(pe0 + 2e') + 4*brev2(pf))).
// Everything after the exchange-A write stays inside the half-warp that shares p1 (tile region 136*p1), as in fft16_core.cuh.
#pragma once
#include "fft16_core.cuh"

namespace tb8 {
using tb::cplx;
using tb::kM;
using tb::kN;
using tb16::brev2;

TB_HD constexpr int brev3(int x) { return ((x & 1) << 2) | (x & 2) | ((x & 4) >> 2); }
// cos/sin(2*pi*e/8)
TB_HD constexpr double w8_cos(int e) { return tb::w32_cos(4 * e); }
TB_HD constexpr double w8_sin(int e) { return tb::w32_sin(4 * e); }

// radix-8 DIF, natural in, bit-reversed out
template <int H>
TB_HD void dif8_stage(double (&re)[8], double (&im)[8]) {
#pragma unroll
    for (int b = 0; b < 8; b += 2 * H) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const int i0 = b + j, i1 = b + j + H;
            const int e = j * (4 / H);   // W_{2H}^j = W8^e
            const double ur = re[i0], ui = im[i0], vr = re[i1], vi = im[i1];
            re[i0] = DADD(ur, vr);
            im[i0] = DADD(ui, vi);
            const double dr = DSUB(ur, vr), di = DSUB(ui, vi);
            if (e == 0) {
                re[i1] = dr; im[i1] = di;
            } else if (e == 2) {          // times -i
                re[i1] = di; im[i1] = -dr;
            } else {
                const double c = w8_cos(e), s = w8_sin(e);
                re[i1] = DFMA(dr, c, DMUL(di, s));
                im[i1] = DFMA(di, c, -DMUL(dr, s));
            }
        }
    }
}
TB_HD void radix8_dif(double (&re)[8], double (&im)[8]) {
    dif8_stage<4>(re, im);
    dif8_stage<2>(re, im);
    dif8_stage<1>(re, im);
}
template <int H>
TB_HD void dit8_stage_inv(double (&re)[8], double (&im)[8]) {
#pragma unroll
    for (int b = 0; b < 8; b += 2 * H) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const int i0 = b + j, i1 = b + j + H;
            const int e = j * (4 / H);
            const double ur = re[i0], ui = im[i0], xr = re[i1], xi = im[i1];
            double vr, vi;
            if (e == 0) { vr = xr; vi = xi; }
            else if (e == 2) { vr = -xi; vi = xr; }
            else {
                const double c = w8_cos(e), s = w8_sin(e);
                vr = DFMA(xr, c, -DMUL(xi, s));
                vi = DFMA(xi, c, DMUL(xr, s));
            }
            re[i0] = DADD(ur, vr);
            im[i0] = DADD(ui, vi);
            re[i1] = DSUB(ur, vr);
            im[i1] = DSUB(ui, vi);
        }
    }
}
TB_HD void radix8_dit_inv(double (&re)[8], double (&im)[8]) {
    dit8_stage_inv<1>(re, im);
    dit8_stage_inv<2>(re, im);
    dit8_stage_inv<4>(re, im);
}

// two radix-4 DIFs on registers [4g .. 4g+3], g = 0, 1 (natural in, bit-reversed out), and the inverse
TB_HD void radix4x2_dif(double (&re)[8], double (&im)[8]) {
#pragma unroll
    for (int g = 0; g < 8; g += 4) {
        const double a0r = DADD(re[g], re[g + 2]), a0i = DADD(im[g], im[g + 2]);
        const double a1r = DADD(re[g + 1], re[g + 3]), a1i = DADD(im[g + 1], im[g + 3]);
        const double d0r = DSUB(re[g], re[g + 2]), d0i = DSUB(im[g], im[g + 2]);
        const double d1r = DSUB(re[g + 1], re[g + 3]), d1i = DSUB(im[g + 1], im[g + 3]);
        re[g] = DADD(a0r, a1r); im[g] = DADD(a0i, a1i);
        re[g + 1] = DSUB(a0r, a1r); im[g + 1] = DSUB(a0i, a1i);
        const double tr = d1i, ti = -d1r;
        re[g + 2] = DADD(d0r, tr); im[g + 2] = DADD(d0i, ti);
        re[g + 3] = DSUB(d0r, tr); im[g + 3] = DSUB(d0i, ti);
    }
}
TB_HD void radix4x2_dit_inv(double (&re)[8], double (&im)[8]) {
#pragma unroll
    for (int g = 0; g < 8; g += 4) {
        const double a0r = DADD(re[g], re[g + 1]), a0i = DADD(im[g], im[g + 1]);
        const double a1r = DSUB(re[g], re[g + 1]), a1i = DSUB(im[g], im[g + 1]);
        const double d0r = DADD(re[g + 2], re[g + 3]), d0i = DADD(im[g + 2], im[g + 3]);
        const double er = DSUB(re[g + 2], re[g + 3]), ei = DSUB(im[g + 2], im[g + 3]);
        const double d1r = -ei, d1i = er;
        re[g] = DADD(a0r, d0r); im[g] = DADD(a0i, d0i);
        re[g + 2] = DSUB(a0r, d0r); im[g + 2] = DSUB(a0i, d0i);
        re[g + 1] = DADD(a1r, d1r); im[g + 1] = DADD(a1i, d1i);
        re[g + 3] = DSUB(a1r, d1r); im[g + 3] = DSUB(a1i, d1i);
    }
}

// z[m] *= exp(+-i*pi*m/16)
TB_HD void pretwist8_fwd(double (&re)[8], double (&im)[8]) {
#pragma unroll
    for (int m = 1; m < 8; ++m) {
        const double c = tb::w32_cos(m), s = tb::w32_sin(m), a = re[m], b = im[m];
        re[m] = DFMA(a, c, -DMUL(b, s));
        im[m] = DFMA(b, c, DMUL(a, s));
    }
}
TB_HD void posttwist8_inv(double (&re)[8], double (&im)[8]) {
#pragma unroll
    for (int m = 1; m < 8; ++m) {
        const double c = tb::w32_cos(m), s = tb::w32_sin(m), a = re[m], b = im[m];
        re[m] = DFMA(a, c, DMUL(b, s));
        im[m] = DFMA(b, c, -DMUL(a, s));
    }
}

// register p *= tw[p] (forward) / conj (inverse); entries flagged trivial (exactly 1) are skipped by the caller's table
template <bool INV>
TB_HD void twiddle8(double (&re)[8], double (&im)[8], const cplx (&tw)[8], int first) {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        if (p < first) continue;
        const double a = re[p], b = im[p];
        if (!INV) {
            re[p] = DFMA(a, tw[p].x, -DMUL(b, tw[p].y));
            im[p] = DFMA(b, tw[p].x, DMUL(a, tw[p].y));
        } else {
            re[p] = DFMA(a, tw[p].x, DMUL(b, tw[p].y));
            im[p] = DFMA(b, tw[p].x, -DMUL(a, tw[p].y));
        }
    }
}

// ---- exchange addressing (16-byte elements; tile = 8 regions of 136 slots, region = p1) --------------------------------------------
constexpr int kTileCplx = 8 * 136;     // 1088 elements, the same 17408 bytes as fft16_core.cuh
// A: writer thread T, register p1: slot 136*p1 + T;   reader thread T' = a + 16*p1, register b: slot 136*p1 + a + 16*b
TB_HD constexpr int xa_wbase(int T) { return T; }
TB_HD constexpr int xa_woff(int p1) { return 136 * p1; }
TB_HD constexpr int xa_rbase(int Tp) { return 136 * (Tp >> 4) + (Tp & 15); }
TB_HD constexpr int xa_roff(int b) { return 16 * b; }
// B: writer thread T' (a = cl + 2c' + 4*d0 + 8*d1), register ps: slot 136*p1 + (cl + 2*d0 + 4*c' + 8*d1) + 17*ps
//    reader thread T'' = c' + 2*ps + 16*p1, register d + 4*cl: slot 136*p1 + 4*c' + 17*ps + (cl + 2*d0 + 8*d1)
TB_HD constexpr int xb_wbase(int Tp) {
    return 136 * (Tp >> 4) + (Tp & 1) + 2 * ((Tp >> 2) & 1) + 4 * ((Tp >> 1) & 1) + 8 * ((Tp >> 3) & 1);
}
TB_HD constexpr int xb_woff(int ps) { return 17 * ps; }
TB_HD constexpr int xb_rbase(int Tpp) { return 136 * (Tpp >> 4) + 4 * (Tpp & 1) + 17 * ((Tpp >> 1) & 7); }
TB_HD constexpr int xb_roff(int r) { return (r >> 2) + 2 * (r & 1) + 8 * ((r >> 1) & 1); }      // r = d + 4*cl
// C: pair (T'' with T'' ^ 1) shares the 16 slots 136*p1 + 17*ps .. +15
//    writer thread c', register pe + 4*cl (pe = pe0 + 2*pe1): slot pairbase + 2*c' + cl + 4*pe0 + 8*pe1
//    reader thread e', register c + 4*pe0 (c = cl + 2*c'):    slot pairbase + 8*e' + c + 4*pe0
TB_HD constexpr int xc_base(int Tpp) { return 136 * (Tpp >> 4) + 17 * ((Tpp >> 1) & 7); }
TB_HD constexpr int xc_wbase(int Tpp) { return xc_base(Tpp) + 2 * (Tpp & 1); }
TB_HD constexpr int xc_woff(int r) { return (r >> 2) + 4 * (r & 1) + 8 * ((r >> 1) & 1); }       // r = pe + 4*cl
TB_HD constexpr int xc_rbase(int Tpp) { return xc_base(Tpp) + 8 * (Tpp & 1); }
TB_HD constexpr int xc_roff(int r) { return r; }                                                  // r = c + 4*pe0

// frequency held by (thread T'' = e' + 2*ps + 16*p1, register pf + 4*pe0) after the forward transform
TB_HD constexpr int freq_of8(int Tpp, int r) {
    return brev3(Tpp >> 4) + 8 * (brev3((Tpp >> 1) & 7) + 8 * (brev2((r >> 2) + 2 * (Tpp & 1)) + 4 * brev2(r & 3)));
}

}  // namespace tb8

// host-side tables (interleaved re, im), one block of 24 complex values per thread T = 0..127:
//   [0..7]   T1[p1] = w^T * W^(T*brev3(p1))                        (pass 1 -> 2; indexed by the thread that holds point T)
//   [8..15]  T2[ps] = W128^(a*brev3(ps)), a = T & 15                (pass 2 -> 3; entry 0 is exactly 1)
//   [16..23] T3[pe + 4*cl] = W16^((cl + 2*(T & 1)) * brev2(pe))     (pass 3 -> 4; entries with pe = 0 are exactly 1)
static inline void tb8_make_tables(double *t /* 2*24*128 */) {
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int T = 0; T < 128; ++T) {
        double *row = t + 2 * 24 * T;
        for (int p = 0; p < 8; ++p) {
            long e = ((long)T * (1 - 4 * (long)tb8::brev3(p))) % 4096;
            if (e < 0) e += 4096;
            row[2 * p] = (double)cosl(pi * (long double)e / 2048.0L);
            row[2 * p + 1] = (double)sinl(pi * (long double)e / 2048.0L);
        }
        for (int ps = 0; ps < 8; ++ps) {
            const int e = ((T & 15) * tb8::brev3(ps)) % 128;
            row[2 * (8 + ps)] = (double)cosl(-2.0L * pi * e / 128.0L);
            row[2 * (8 + ps) + 1] = (double)sinl(-2.0L * pi * e / 128.0L);
        }
        for (int r = 0; r < 8; ++r) {
            const int cl = r >> 2, pe = r & 3, e = ((cl + 2 * (T & 1)) * tb16::brev2(pe)) % 16;
            row[2 * (16 + r)] = (double)cosl(-2.0L * pi * e / 16.0L);
            row[2 * (16 + r) + 1] = (double)sinl(-2.0L * pi * e / 16.0L);
        }
    }
}
