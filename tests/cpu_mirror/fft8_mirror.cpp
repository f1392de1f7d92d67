// Thread-by-thread CPU emulation of the 128-thread x 8-point FFT in fhe_string_bounty_b200/csrc/fft8_core.cuh: forward transform
// against the definition Z_k = sum_j z_j w^j W^(jk), the (thread, register) -> frequency map, inverse(forward(x)) == 1024 x, injective
// exchange addressing, region confinement of everything after the exchange-A write, and the bank behaviour of each exchange side.
#include "../../fhe_string_bounty_b200/csrc/fft8_core.cuh"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

using namespace tb8;
typedef std::complex<long double> cld;

struct Poly { double re[128][8], im[128][8]; };

static bool exchange(Poly &w, int (*wr)(int, int), int (*rd)(int, int), bool inverse) {
    static tb::cplx tile[kTileCplx];
    std::vector<int> hit(kTileCplx, 0);
    for (int t = 0; t < 128; ++t)
        for (int p = 0; p < 8; ++p) {
            const int a = inverse ? rd(t, p) : wr(t, p);
            if (a < 0 || a >= kTileCplx) return false;
            ++hit[a];
            tile[a].x = w.re[t][p]; tile[a].y = w.im[t][p];
        }
    for (int a = 0; a < kTileCplx; ++a) if (hit[a] > 1) return false;
    for (int t = 0; t < 128; ++t)
        for (int p = 0; p < 8; ++p) {
            const int a = inverse ? wr(t, p) : rd(t, p);
            w.re[t][p] = tile[a].x; w.im[t][p] = tile[a].y;
        }
    return true;
}
static int xaw(int t, int p) { return xa_wbase(t) + xa_woff(p); }
static int xar(int t, int p) { return xa_rbase(t) + xa_roff(p); }
static int xbw(int t, int p) { return xb_wbase(t) + xb_woff(p); }
static int xbr(int t, int p) { return xb_rbase(t) + xb_roff(p); }
static int xcw(int t, int p) { return xc_wbase(t) + xc_woff(p); }
static int xcr(int t, int p) { return xc_rbase(t) + xc_roff(p); }

// worst number of lanes of a quarter-warp (8 lanes x 16 bytes) that fall into the same 16-byte bank
static int worst_conflict(int (*f)(int, int)) {
    int worst = 0;
    for (int p = 0; p < 8; ++p)
        for (int q = 0; q < 16; ++q) {
            int cnt[8] = {0};
            for (int l = 0; l < 8; ++l) ++cnt[f(8 * q + l, p) & 7];
            for (int b = 0; b < 8; ++b) if (cnt[b] > worst) worst = cnt[b];
        }
    return worst;
}

int main() {
    std::vector<tb::cplx> tab(24 * 128);
    tb8_make_tables(reinterpret_cast<double *>(tab.data()));
    const long double pi = 3.14159265358979323846264338327950288L;

    if (worst_conflict(xaw) != 1 || worst_conflict(xar) != 1 || worst_conflict(xbw) != 1 || worst_conflict(xbr) != 1) { printf("FAIL bank conflicts A/B\n"); return 1; }
    printf("exchange C: worst %d-way (write), %d-way (read)\n", worst_conflict(xcw), worst_conflict(xcr));
    if (worst_conflict(xcw) > 2 || worst_conflict(xcr) > 2) { printf("FAIL bank conflicts C\n"); return 1; }
    for (int t = 0; t < 128; ++t)
        for (int g = 0; g < 8; ++g)
            for (int a : {xar(t, g), xbw(t, g), xbr(t, g), xcw(t, g), xcr(t, g)})
                if (a / 136 != (t >> 4)) { printf("FAIL region confinement\n"); return 1; }

    srand(13);
    std::vector<cld> z(kM);
    Poly w;
    for (int j = 0; j < kM; ++j) {
        double a = (double)((rand() % (1 << 23)) - (1 << 22)), b = (double)((rand() % (1 << 23)) - (1 << 22));
        z[j] = cld(a, b);
        w.re[j & 127][j >> 7] = a;   // thread T = j mod 128, register m = j / 128
        w.im[j & 127][j >> 7] = b;
    }
    Poly orig = w;
    auto tw = [&](int t, int block, tb::cplx (&out)[8]) { for (int p = 0; p < 8; ++p) out[p] = tab[24 * t + 8 * block + p]; };

    for (int t = 0; t < 128; ++t) {
        tb::cplx c[8]; tw(t, 0, c);
        pretwist8_fwd(w.re[t], w.im[t]); radix8_dif(w.re[t], w.im[t]); twiddle8<false>(w.re[t], w.im[t], c, 0);
    }
    if (!exchange(w, xaw, xar, false)) { printf("FAIL exchange A\n"); return 1; }
    for (int t = 0; t < 128; ++t) {
        tb::cplx c[8]; tw(t, 1, c);
        radix8_dif(w.re[t], w.im[t]); twiddle8<false>(w.re[t], w.im[t], c, 1);
    }
    if (!exchange(w, xbw, xbr, false)) { printf("FAIL exchange B\n"); return 1; }
    for (int t = 0; t < 128; ++t) {
        tb::cplx c[8]; tw(t, 2, c);
        radix4x2_dif(w.re[t], w.im[t]); twiddle8<false>(w.re[t], w.im[t], c, 0);
    }
    if (!exchange(w, xcw, xcr, false)) { printf("FAIL exchange C\n"); return 1; }
    for (int t = 0; t < 128; ++t) radix4x2_dif(w.re[t], w.im[t]);

    long double max_err = 0, max_mag = 0;
    std::set<int> seen;
    for (int t = 0; t < 128; ++t)
        for (int g = 0; g < 8; ++g) {
            const int k = freq_of8(t, g);
            seen.insert(k);
            cld acc = 0;
            for (int j = 0; j < kM; ++j) {
                long e = ((long)j * (1 - 4 * (long)k)) % 4096;
                acc += z[j] * std::polar(1.0L, pi * (long double)e / 2048.0L);
            }
            long double err = std::abs(acc - cld(w.re[t][g], w.im[t][g]));
            if (err > max_err) max_err = err;
            if (std::abs(acc) > max_mag) max_mag = std::abs(acc);
        }
    printf("forward: max |err| = %.3Le (max |Z| = %.3Le, rel %.3Le)\n", max_err, max_mag, max_err / max_mag);
    if (seen.size() != 1024) { printf("FAIL frequency map not a bijection\n"); return 1; }
    if (max_err / max_mag > 1e-14L) { printf("FAIL forward\n"); return 1; }

    for (int t = 0; t < 128; ++t) radix4x2_dit_inv(w.re[t], w.im[t]);
    if (!exchange(w, xcw, xcr, true)) { printf("FAIL exchange C inverse\n"); return 1; }
    for (int t = 0; t < 128; ++t) {
        tb::cplx c[8]; tw(t, 2, c);
        twiddle8<true>(w.re[t], w.im[t], c, 0); radix4x2_dit_inv(w.re[t], w.im[t]);
    }
    if (!exchange(w, xbw, xbr, true)) { printf("FAIL exchange B inverse\n"); return 1; }
    for (int t = 0; t < 128; ++t) {
        tb::cplx c[8]; tw(t, 1, c);
        twiddle8<true>(w.re[t], w.im[t], c, 1); radix8_dit_inv(w.re[t], w.im[t]);
    }
    if (!exchange(w, xaw, xar, true)) { printf("FAIL exchange A inverse\n"); return 1; }
    for (int t = 0; t < 128; ++t) {
        tb::cplx c[8]; tw(t, 0, c);
        twiddle8<true>(w.re[t], w.im[t], c, 0); radix8_dit_inv(w.re[t], w.im[t]); posttwist8_inv(w.re[t], w.im[t]);
    }
    long double max_rt = 0;
    for (int t = 0; t < 128; ++t)
        for (int m = 0; m < 8; ++m) {
            long double er = fabsl((long double)w.re[t][m] / 1024.0L - orig.re[t][m]);
            long double ei = fabsl((long double)w.im[t][m] / 1024.0L - orig.im[t][m]);
            if (er > max_rt) max_rt = er;
            if (ei > max_rt) max_rt = ei;
        }
    printf("roundtrip: max |err| = %.3Le on inputs of magnitude 2^22\n", max_rt);
    if (max_rt > 1e-8L) { printf("FAIL roundtrip\n"); return 1; }
    printf("OK\n");
    return 0;
}
