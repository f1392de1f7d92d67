// Thread-by-thread CPU emulation of the 64-thread x 16-point FFT in fhe_string_bounty_b200/csrc/fft16_core.cuh.
// Checks the forward transform against the definition Z_k = sum_j z_j w^j W^(jk), the (thread, register) -> frequency map,
// inverse(forward(x)) == 1024 x, that both exchanges map injectively into the padded tile, and that every quarter-warp
// access (8 lanes x 16 bytes) of both exchanges touches 8 distinct 16-byte banks.
#include "../../fhe_string_bounty_b200/csrc/fft16_core.cuh"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

using namespace tb16;
typedef std::complex<long double> cld;

struct Poly {
    double re[64][16], im[64][16];   // [thread][register]
};

static bool exchange(Poly &w, int (*wr)(int, int), int (*rd)(int, int), bool inverse) {
    static tb::cplx tile[kTileCplx];
    std::vector<int> hit(kTileCplx, 0);
    // forward: thread t writes register p at wr(t, p), then reads register u from rd(t, u); inverse: swap roles
    for (int t = 0; t < 64; ++t)
        for (int p = 0; p < 16; ++p) {
            const int a = inverse ? rd(t, p) : wr(t, p);
            if (a < 0 || a >= kTileCplx) return false;
            ++hit[a];
            tile[a].x = w.re[t][p]; tile[a].y = w.im[t][p];
        }
    for (int a = 0; a < kTileCplx; ++a) if (hit[a] > 1) return false;   // injective
    for (int t = 0; t < 64; ++t)
        for (int p = 0; p < 16; ++p) {
            const int a = inverse ? wr(t, p) : rd(t, p);
            w.re[t][p] = tile[a].x; w.im[t][p] = tile[a].y;
        }
    return true;
}
static int xaw(int t, int p) { return xa_write(t, p); }
static int xar(int t, int p) { return xa_read(t, p); }
static int xbw(int t, int p) { return xb_write(t, p); }
static int xbr(int t, int p) { return xb_read(t, p); }

static bool conflict_free(int (*f)(int, int)) {
    for (int p = 0; p < 16; ++p)
        for (int q = 0; q < 8; ++q) {      // 8 quarter-warps of the 64 threads
            std::set<int> banks;
            for (int l = 0; l < 8; ++l) banks.insert(f(8 * q + l, p) & 7);
            if (banks.size() != 8) return false;
        }
    return true;
}

int main() {
    std::vector<tb::cplx> t1(1024), t2(64);
    tb16_make_tables(reinterpret_cast<double *>(t1.data()), reinterpret_cast<double *>(t2.data()));
    const long double pi = 3.14159265358979323846264338327950288L;

    if (!conflict_free(xaw) || !conflict_free(xar) || !conflict_free(xbw) || !conflict_free(xbr)) { printf("FAIL bank conflicts\n"); return 1; }

    // everything after the exchange-A write stays inside the region of the half-warp that shares kq = thread >> 4
    for (int t = 0; t < 64; ++t)
        for (int g = 0; g < 16; ++g)
            for (int a : {xa_read(t, g), xb_write(t, g), xb_read(t, g)})
                if (a / 272 != (t >> 4)) { printf("FAIL region confinement\n"); return 1; }

    srand(11);
    std::vector<cld> z(kM);
    Poly w;
    for (int j = 0; j < kM; ++j) {
        double a = (double)((rand() % (1 << 23)) - (1 << 22)), b = (double)((rand() % (1 << 23)) - (1 << 22));
        z[j] = cld(a, b);
        w.re[j & 63][j >> 6] = a;   // thread T = j mod 64, register m = j / 64
        w.im[j & 63][j >> 6] = b;
    }
    Poly orig = w;

    // forward
    for (int t = 0; t < 64; ++t) {
        radix16_twisted_fwd(w.re[t], w.im[t]);
        twiddle16_fwd(w.re[t], w.im[t], [&](int p) { return t1[p * 64 + t]; });
    }
    if (!exchange(w, xaw, xar, false)) { printf("FAIL exchange A\n"); return 1; }
    for (int t = 0; t < 64; ++t) {
        const tb::cplx tw[3] = {t2[t & 15], t2[16 + (t & 15)], t2[32 + (t & 15)]};
        radix4x4_dif(w.re[t], w.im[t]);
        twiddle4_fwd(w.re[t], w.im[t], tw);
    }
    if (!exchange(w, xbw, xbr, false)) { printf("FAIL exchange B\n"); return 1; }
    for (int t = 0; t < 64; ++t) radix16_fwd(w.re[t], w.im[t]);

    long double max_err = 0, max_mag = 0;
    std::set<int> seen;
    for (int t = 0; t < 64; ++t)
        for (int g = 0; g < 16; ++g) {
            const int k = freq_of16(t, g);
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

    // inverse
    for (int t = 0; t < 64; ++t) radix16_dit_inv(w.re[t], w.im[t]);
    if (!exchange(w, xbw, xbr, true)) { printf("FAIL exchange B inverse\n"); return 1; }
    for (int t = 0; t < 64; ++t) {
        const tb::cplx tw[3] = {t2[t & 15], t2[16 + (t & 15)], t2[32 + (t & 15)]};
        twiddle4_inv(w.re[t], w.im[t], tw);
        radix4x4_dit_inv(w.re[t], w.im[t]);
    }
    if (!exchange(w, xaw, xar, true)) { printf("FAIL exchange A inverse\n"); return 1; }
    for (int t = 0; t < 64; ++t) {
        twiddle16_inv(w.re[t], w.im[t], [&](int p) { return t1[p * 64 + t]; });
        radix16_dit_inv(w.re[t], w.im[t]);
        posttwist16_inv(w.re[t], w.im[t]);
    }
    long double max_rt = 0;
    for (int t = 0; t < 64; ++t)
        for (int m = 0; m < 16; ++m) {
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
