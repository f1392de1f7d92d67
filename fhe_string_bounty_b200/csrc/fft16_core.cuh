// fft16_core.cuh -- the same size-1024 negacyclic-carrier FFT as fft_core.cuh, laid out for 64 threads x 16 points
// (two warps per polynomial, ~half the registers per thread, twice the warps per SM).
//
// Point j = T + 64*m (thread T = 0..63, register m = 0..15).  Z_k = sum_j z_j w^j W^(jk), w = exp(i*pi/2048),
// W = exp(-2*pi*i/1024), k = k1 + 16*k2, k2 = s + 4*v, T = u + 16*r:
//   pass 1   in-register radix-16 pass over m of z * w^(64 m) (radix16_twisted_fwd)  -> register p1 = brev4(k1)
//   twiddle  T1[p1][T] = w^T * W^(T * brev4(p1))
//   exchange (T = u + 16r, p1 = 4*kq + pl) -> thread T' = u + 16*kq, register 4*pl + r     (64-point DFT over T remains)
//   pass 2   four radix-4 DIFs over r                                                -> register 4*pl + ps, s = brev2(ps)
//   twiddle  W64^(u * s): THREE per-thread constants (u = T' & 15), nothing register dependent
//   exchange (T' = u + 16*kq, 4*pl + ps) -> thread T'' = ps + 4*(4*kq + pl), register u
//   pass 3   radix-16 pass over u (radix16_fwd)                                      -> register pv = brev4(v)
// Thread T'' = ps + 4*p1, register pv holds k = brev4(p1) + 16*(brev2(ps) + 4*brev4(pv)).
// The inverse runs the same steps backwards with conjugate twiddles.  Shared-memory exchanges use 16-byte elements in
// one tile per polynomial (conflict free both ways, see xa_* / xb_*).  Compiles for host and device like fft_core.cuh.
#pragma once
#include "fft_core.cuh"

namespace tb16 {
using tb::cplx;
using tb::kM;
using tb::kN;

TB_HD constexpr int brev4(int x) { return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3); }
TB_HD constexpr int brev2(int x) { return ((x & 1) << 1) | ((x & 2) >> 1); }

// cos/sin(2*pi*e/16), e = 0..7 (W16^e = c - i s)
TB_HD constexpr double w16_cos(int e) { return tb::w32_cos(2 * e); }
TB_HD constexpr double w16_sin(int e) { return tb::w32_sin(2 * e); }
// pre-twist w^(64 m) = exp(i*pi*m/32), m = 0..15  == fft_core's exp(i*pi*(2m)/64)
TB_HD constexpr double pre16_cos(int m) { return tb::pre_cos(2 * m); }
TB_HD constexpr double pre16_sin(int m) { return tb::pre_sin(2 * m); }

// cos / sin(pi*e/32) for any integer e (compile-time after unrolling)
TB_HD constexpr double cis32_cos(int e) {
    const int x = 2 * (((e % 64) + 64) % 64);     // angle pi*x/64, x in [0, 128)
    return x < 32 ? tb::pre_cos(x) : x == 32 ? 0.0 : x <= 64 ? -tb::pre_cos(64 - x) : x < 96 ? -tb::pre_cos(x - 64) : x == 96 ? 0.0 : tb::pre_cos(128 - x);
}
TB_HD constexpr double cis32_sin(int e) { return cis32_cos(e - 16); }

// ---- radix-16 forward pass on registers [0 .. 15]: natural in, bit-reversed out ---------------------------------------------------
// Evaluation of Z(X) = sum_m z_m X^m at x_k = exp(i*pi*(t - 4k)/32), k = 0..15, result for k in register brev4(k); t = 0: the plain
// size-16 DFT (kernel exp(-2*pi*i*m*k/16)), t = 1 (TWISTED): the DFT of z_m * exp(i*pi*m/32), i.e. the negacyclic pre-twist folded in.
// Stage H reduces a block of 2H registers modulo X^H - C and X^H + C, C = x_k^H for the block's k (the low bits of k fixed so far,
// k_rep = brev4(block base)): (a, b) -> (a + C b, a - C b) with the twiddle BEFORE the add, so a non-trivial butterfly is six FMAs in
// the tangent form (p = b_r - t b_i, q = b_i + t b_r; a +- c (p, q)) instead of the four adds + four multiply-adds of the
// twiddle-after-subtract form, and the pre-twist costs nothing: 148 (plain) / 192 (twisted) FP64 instructions against 168 / 228.
template <int H, bool TWISTED>
TB_HD void radix16_stage_fwd(double (&re)[16], double (&im)[16]) {
#pragma unroll
    for (int b = 0; b < 16; b += 2 * H) {
        const int e = (((TWISTED ? 1 : 0) - 4 * brev4(b)) * H % 64 + 64) % 64;       // C = exp(i*pi*e/32)
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const int i0 = b + j, i1 = b + j + H;
            const double ur = re[i0], ui = im[i0], xr = re[i1], xi = im[i1];
            if (e % 16 == 0) {           // C = 1, i, -1, -i
                double vr, vi;
                if (e == 0) { vr = xr; vi = xi; } else if (e == 16) { vr = -xi; vi = xr; } else if (e == 32) { vr = -xr; vi = -xi; } else { vr = xi; vi = -xr; }
                re[i0] = DADD(ur, vr);
                im[i0] = DADD(ui, vi);
                re[i1] = DSUB(ur, vr);
                im[i1] = DSUB(ui, vi);
            } else {
                const double c = cis32_cos(e), t = cis32_sin(e) / cis32_cos(e);
                const double p = DFMA(-t, xi, xr);
                const double q = DFMA(t, xr, xi);
                re[i0] = DFMA(c, p, ur);
                im[i0] = DFMA(c, q, ui);
                re[i1] = DFMA(-c, p, ur);
                im[i1] = DFMA(-c, q, ui);
            }
        }
    }
}
TB_HD void radix16_fwd(double (&re)[16], double (&im)[16]) {
    radix16_stage_fwd<8, false>(re, im);
    radix16_stage_fwd<4, false>(re, im);
    radix16_stage_fwd<2, false>(re, im);
    radix16_stage_fwd<1, false>(re, im);
}
// pre-twist w^(64 m) = exp(i*pi*m/32) + radix-16 forward pass in one
TB_HD void radix16_twisted_fwd(double (&re)[16], double (&im)[16]) {
    radix16_stage_fwd<8, true>(re, im);
    radix16_stage_fwd<4, true>(re, im);
    radix16_stage_fwd<2, true>(re, im);
    radix16_stage_fwd<1, true>(re, im);
}

// exact inverse up to a factor 16 (DIT, conjugate twiddles, cosine factored out as in fft_core.cuh)
template <int H>
TB_HD void dit16_stage_inv(double (&re)[16], double (&im)[16]) {
#pragma unroll
    for (int b = 0; b < 16; b += 2 * H) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const int i0 = b + j, i1 = b + j + H;
            const int e = j * (8 / H);
            const double ur = re[i0], ui = im[i0], xr = re[i1], xi = im[i1];
            if (e == 0 || e == 4) {
                double vr, vi;
                if (e == 0) { vr = xr; vi = xi; } else { vr = -xi; vi = xr; }
                re[i0] = DADD(ur, vr);
                im[i0] = DADD(ui, vi);
                re[i1] = DSUB(ur, vr);
                im[i1] = DSUB(ui, vi);
            } else {
                const double c = w16_cos(e), t = w16_sin(e) / w16_cos(e);
                const double p = DFMA(-t, xi, xr);
                const double q = DFMA(t, xr, xi);
                re[i0] = DFMA(c, p, ur);
                im[i0] = DFMA(c, q, ui);
                re[i1] = DFMA(-c, p, ur);
                im[i1] = DFMA(-c, q, ui);
            }
        }
    }
}
TB_HD void radix16_dit_inv(double (&re)[16], double (&im)[16]) {
    dit16_stage_inv<1>(re, im);
    dit16_stage_inv<2>(re, im);
    dit16_stage_inv<4>(re, im);
    dit16_stage_inv<8>(re, im);
}

// ---- four radix-4 DIFs on register groups [4g .. 4g+3] (natural in, bit-reversed out) and their inverse ---------------------------
TB_HD void radix4x4_dif(double (&re)[16], double (&im)[16]) {
#pragma unroll
    for (int g = 0; g < 16; g += 4) {
        const double a0r = DADD(re[g], re[g + 2]), a0i = DADD(im[g], im[g + 2]);
        const double a1r = DADD(re[g + 1], re[g + 3]), a1i = DADD(im[g + 1], im[g + 3]);
        const double d0r = DSUB(re[g], re[g + 2]), d0i = DSUB(im[g], im[g + 2]);
        const double d1r = DSUB(re[g + 1], re[g + 3]), d1i = DSUB(im[g + 1], im[g + 3]);
        // stage 2: (a0, a1) -> positions 0, 1 ; (d0, -i*d1) -> positions 2, 3
        re[g] = DADD(a0r, a1r); im[g] = DADD(a0i, a1i);
        re[g + 1] = DSUB(a0r, a1r); im[g + 1] = DSUB(a0i, a1i);
        const double tr = d1i, ti = -d1r;     // d1 * (-i)
        re[g + 2] = DADD(d0r, tr); im[g + 2] = DADD(d0i, ti);
        re[g + 3] = DSUB(d0r, tr); im[g + 3] = DSUB(d0i, ti);
    }
}
TB_HD void radix4x4_dit_inv(double (&re)[16], double (&im)[16]) {
#pragma unroll
    for (int g = 0; g < 16; g += 4) {
        const double a0r = DADD(re[g], re[g + 1]), a0i = DADD(im[g], im[g + 1]);
        const double a1r = DSUB(re[g], re[g + 1]), a1i = DSUB(im[g], im[g + 1]);
        const double d0r = DADD(re[g + 2], re[g + 3]), d0i = DADD(im[g + 2], im[g + 3]);
        const double er = DSUB(re[g + 2], re[g + 3]), ei = DSUB(im[g + 2], im[g + 3]);
        const double d1r = -ei, d1i = er;     // e * (+i)
        re[g] = DADD(a0r, d0r); im[g] = DADD(a0i, d0i);
        re[g + 2] = DSUB(a0r, d0r); im[g + 2] = DSUB(a0i, d0i);
        re[g + 1] = DADD(a1r, d1r); im[g + 1] = DADD(a1i, d1i);
        re[g + 3] = DSUB(a1r, d1r); im[g + 3] = DSUB(a1i, d1i);
    }
}

TB_HD void posttwist16_inv(double (&re)[16], double (&im)[16]) {
#pragma unroll
    for (int m = 1; m < 16; ++m) {
        const double a = re[m], b = im[m];
        if (m == 8) {
            const double c = pre16_cos(m);
            re[m] = DMUL(DADD(a, b), c);
            im[m] = DMUL(DSUB(b, a), c);
        } else {            // (a + i b) * (c - i s) = c * ((a + t b) + i (b - t a))
            const double c = pre16_cos(m), t = pre16_sin(m) / pre16_cos(m);
            re[m] = DMUL(c, DFMA(t, b, a));
            im[m] = DMUL(c, DFMA(-t, a, b));
        }
    }
}

// register p of thread T *= load(p*64 + T)  (forward) / conj (inverse): table T1[p1][T] (1024 entries) or T2[p2][r] via load
template <class Load>
TB_HD void twiddle16_fwd(double (&re)[16], double (&im)[16], Load load) {
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        const cplx w = load(p);
        const double a = re[p], b = im[p];
        re[p] = DFMA(a, w.x, -DMUL(b, w.y));
        im[p] = DFMA(b, w.x, DMUL(a, w.y));
    }
}
template <class Load>
TB_HD void twiddle16_inv(double (&re)[16], double (&im)[16], Load load) {
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        const cplx w = load(p);
        const double a = re[p], b = im[p];
        re[p] = DFMA(a, w.x, DMUL(b, w.y));
        im[p] = DFMA(b, w.x, -DMUL(a, w.y));
    }
}

// the three per-thread twiddles of pass 2 -> 3, applied to register 4*pl + ps (ps = 1, 2, 3): tw[ps - 1] = W64^(u * brev2(ps))
TB_HD void twiddle4_fwd(double (&re)[16], double (&im)[16], const cplx (&tw)[3]) {
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        if ((g & 3) == 0) continue;
        const cplx w = tw[(g & 3) - 1];
        const double a = re[g], b = im[g];
        re[g] = DFMA(a, w.x, -DMUL(b, w.y));
        im[g] = DFMA(b, w.x, DMUL(a, w.y));
    }
}
TB_HD void twiddle4_inv(double (&re)[16], double (&im)[16], const cplx (&tw)[3]) {
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        if ((g & 3) == 0) continue;
        const cplx w = tw[(g & 3) - 1];
        const double a = re[g], b = im[g];
        re[g] = DFMA(a, w.x, DMUL(b, w.y));
        im[g] = DFMA(b, w.x, -DMUL(a, w.y));
    }
}

// ---- exchange addressing (units of 16-byte complex elements; one tile of kTileCplx elements per polynomial) --------------------------
// Every address is "per-thread base + compile-time constant" (nothing to keep alive in registers across the blind-rotation loop) and
// every quarter-warp access (8 lanes x 16 bytes) touches 8 distinct 16-byte banks on both sides of both exchanges:
// The tile is split into four regions of 272 slots, one per kq: everything after the exchange-A write stays inside the 16 threads
// (one half-warp) that share kq, so only exchange A needs a barrier across the polynomial's two warps.
// exchange A: (thread T = u + 16r, register p1 = 4*kq + pl) <-> (thread T' = u + 16*kq, register 4*pl + r): slot 272kq + u + 16r + 64pl
//   writer: [T] + 272*(p1 >> 2) + 64*(p1 & 3)       reader: [272*kq + u] + 64*(g >> 2) + 16*(g & 3)
// exchange B: (thread T' = u + 16*kq, register g = 4*pl + ps) <-> (thread T'' = ps + 4*pl + 16*kq, register u): slot 272kq + 17u + g
//   writer: [272*kq + 17*u] + g                     reader: [272*kq + (T'' & 15)] + 17*u
// The reader side of B is also where a thread parks its 16 spectrum values for the partner polynomial's thread of the same index.
constexpr int kTileCplx = 4 * 272;     // 1088 elements = 17408 bytes
TB_HD constexpr int xa_wbase(int T) { return T; }
TB_HD constexpr int xa_woff(int p1) { return 272 * (p1 >> 2) + 64 * (p1 & 3); }
TB_HD constexpr int xa_rbase(int Tp) { return 272 * (Tp >> 4) + (Tp & 15); }
TB_HD constexpr int xa_roff(int g) { return 64 * (g >> 2) + 16 * (g & 3); }
TB_HD constexpr int xb_wbase(int Tp) { return 272 * (Tp >> 4) + 17 * (Tp & 15); }
TB_HD constexpr int xb_woff(int g) { return g; }
TB_HD constexpr int xb_rbase(int Tpp) { return 272 * (Tpp >> 4) + (Tpp & 15); }
TB_HD constexpr int xb_roff(int u) { return 17 * u; }
TB_HD constexpr int xa_write(int T, int p1) { return xa_wbase(T) + xa_woff(p1); }
TB_HD constexpr int xa_read(int Tp, int g) { return xa_rbase(Tp) + xa_roff(g); }
TB_HD constexpr int xb_write(int Tp, int g) { return xb_wbase(Tp) + xb_woff(g); }
TB_HD constexpr int xb_read(int Tpp, int u) { return xb_rbase(Tpp) + xb_roff(u); }

// frequency held by (thread T'' = ps + 4*p1, register pv) after the forward transform
TB_HD constexpr int freq_of16(int Tpp, int pv) { return brev4(Tpp >> 2) + 16 * (brev2(Tpp & 3) + 4 * brev4(pv)); }

}  // namespace tb16

// host-side tables: T1[p1*64 + T] = w^T * W^(T*brev4(p1)) = exp(i*pi*T*(1 - 4*brev4(p1))/2048);
// T2[(ps - 1)*16 + u] = W64^(u*brev2(ps)), ps = 1..3 (48 entries, stored in 64)
static inline void tb16_make_tables(double *t1 /* 2*1024 */, double *t2 /* 2*64 */) {
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int p = 0; p < 16; ++p)
        for (int T = 0; T < 64; ++T) {
            long e = ((long)T * (1 - 4 * (long)tb16::brev4(p))) % 4096;
            if (e < 0) e += 4096;
            t1[2 * (p * 64 + T)] = (double)cosl(pi * (long double)e / 2048.0L);
            t1[2 * (p * 64 + T) + 1] = (double)sinl(pi * (long double)e / 2048.0L);
        }
    for (int i = 0; i < 128; ++i) t2[i] = 0.0;
    for (int ps = 1; ps < 4; ++ps)
        for (int u = 0; u < 16; ++u) {
            const int e = (u * tb16::brev2(ps)) % 64;
            t2[2 * ((ps - 1) * 16 + u)] = (double)cosl(-2.0L * pi * e / 64.0L);
            t2[2 * ((ps - 1) * 16 + u) + 1] = (double)sinl(-2.0L * pi * e / 64.0L);
        }
}
