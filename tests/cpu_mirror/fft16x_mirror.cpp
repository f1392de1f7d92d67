// Thread-by-thread CPU emulation of the two 16-points-per-thread FFTs added in round 2, from the same header the kernels use
// (fhe_string_bounty_b200/csrc/fft16x_slots.cuh + the radix-16 butterflies of fft16_core.cuh):
//   * 256 points, 16 threads (pbs_n512.cu):   radix 16 -> exchange -> radix 16
//   * 4096 points, 256 threads (pbs_n8192.cu): radix 16 -> exchange A -> radix 16 -> exchange B -> radix 16
// Checks: forward == the definition Z_k = sum_j z_j w^j W^(jk) at the documented (thread, register) -> frequency map; inverse(forward(x))
// == M x; every exchange maps injectively into its tile; every quarter-warp access (8 lanes x 16 bytes) touches 8 distinct 16-byte banks;
// in the 4096-point transform everything after the exchange-A write stays inside the half-warp's region.
#include "../../fhe_string_bounty_b200/csrc/fft16x_slots.cuh"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

using namespace tb16;
using namespace tb16x;
typedef std::complex<long double> cld;
static const long double PI = 3.14159265358979323846264338327950288L;

struct Regs { double re[16], im[16]; };

// thread t writes register p at wr(t, p), then reads register u from rd(t, u)  (inverse: roles swapped)
template <class W, class R>
static bool exchange(std::vector<Regs> &th, int tile_size, W wr, R rd, bool inverse) {
    std::vector<tb::cplx> tile(tile_size);
    std::vector<int> hit(tile_size, 0);
    for (size_t t = 0; t < th.size(); ++t)
        for (int p = 0; p < 16; ++p) {
            const int a = inverse ? rd((int)t, p) : wr((int)t, p);
            if (a < 0 || a >= tile_size) return false;
            if (++hit[a] > 1) return false;
            tile[a].x = th[t].re[p]; tile[a].y = th[t].im[p];
        }
    for (size_t t = 0; t < th.size(); ++t)
        for (int p = 0; p < 16; ++p) {
            const int a = inverse ? wr((int)t, p) : rd((int)t, p);
            th[t].re[p] = tile[a].x; th[t].im[p] = tile[a].y;
        }
    return true;
}
template <class F>
static bool conflict_free(int n_threads, F f) {
    for (int p = 0; p < 16; ++p)
        for (int q = 0; q < n_threads / 8; ++q) {
            std::set<int> banks;
            for (int l = 0; l < 8; ++l) banks.insert(f(8 * q + l, p) & 7);
            if (banks.size() != 8) return false;
        }
    return true;
}
static void mul(Regs &r, int p, tb::cplx w, bool conj) {
    const double a = r.re[p], b = r.im[p], s = conj ? -w.y : w.y;
    r.re[p] = a * w.x - b * s; r.im[p] = b * w.x + a * s;
}

static int check(int M, int n_threads) {
    const int N = 2 * M;
    std::vector<tb::cplx> tbl(M == 256 ? 256 : 4352);
    if (M == 256) tb16x_make_table_512(reinterpret_cast<double *>(tbl.data())); else tb16x_make_table_8192(reinterpret_cast<double *>(tbl.data()));
    srand(17 + M);
    std::vector<cld> z(M);
    std::vector<Regs> th(n_threads);
    for (int j = 0; j < M; ++j) {
        const double a = (double)((rand() % (1 << 18)) - (1 << 17)), b = (double)((rand() % (1 << 18)) - (1 << 17));
        z[j] = cld(a, b);
        th[j % n_threads].re[j / n_threads] = a;      // point j = T + n_threads * m
        th[j % n_threads].im[j / n_threads] = b;
    }
    const std::vector<Regs> orig = th;
    auto freq = [&](int t, int r) { return M == 256 ? freq256(t, r) : freq4096(t, r); };

    // ---- forward ----
    for (int t = 0; t < n_threads; ++t) {
        radix16_twisted_fwd(th[t].re, th[t].im);
        for (int p = 0; p < 16; ++p) mul(th[t], p, tbl[p * n_threads + t], false);
    }
    if (M == 256) {
        if (!exchange(th, kTile256, s256_write, s256_read, false)) { printf("FAIL 256 exchange\n"); return 1; }
        for (int t = 0; t < n_threads; ++t) radix16_fwd(th[t].re, th[t].im);
    } else {
        if (!exchange(th, kTile4096, s4096_a_write, s4096_a_read, false)) { printf("FAIL 4096 exchange A\n"); return 1; }
        for (int t = 0; t < n_threads; ++t) {
            radix16_fwd(th[t].re, th[t].im);
            for (int p = 0; p < 16; ++p) mul(th[t], p, tbl[4096 + p * 16 + (t & 15)], false);
        }
        if (!exchange(th, kTile4096, s4096_b_write, s4096_b_read, false)) { printf("FAIL 4096 exchange B\n"); return 1; }
        for (int t = 0; t < n_threads; ++t) radix16_fwd(th[t].re, th[t].im);
    }
    // against the definition, on a sample of frequencies (the full check is O(M^2) = fine for 256, sampled for 4096)
    std::set<int> seen;
    long double worst = 0;
    for (int t = 0; t < n_threads; ++t)
        for (int r = 0; r < 16; ++r) {
            const int k = freq(t, r);
            if (k < 0 || k >= M || !seen.insert(k).second) { printf("FAIL frequency map is not a permutation (M = %d)\n", M); return 1; }
            if (M == 4096 && ((t * 16 + r) % 37) != 0) continue;
            cld acc = 0;
            for (int j = 0; j < M; ++j) {
                const long double ang = PI * j / N - 2.0L * PI * (long double)((long long)j * k % M) / M;
                acc += z[j] * cld(cosl(ang), sinl(ang));
            }
            worst = std::max(worst, std::abs(acc - cld(th[t].re[r], th[t].im[r])));
        }
    if (worst > 1e-6L * M) { printf("FAIL forward transform (M = %d): max error %Lg\n", M, worst); return 1; }

    // ---- inverse ----
    for (int t = 0; t < n_threads; ++t) radix16_dit_inv(th[t].re, th[t].im);
    if (M == 256) {
        if (!exchange(th, kTile256, s256_write, s256_read, true)) { printf("FAIL 256 inverse exchange\n"); return 1; }
    } else {
        if (!exchange(th, kTile4096, s4096_b_write, s4096_b_read, true)) { printf("FAIL 4096 inverse exchange B\n"); return 1; }
        for (int t = 0; t < n_threads; ++t) {
            for (int p = 0; p < 16; ++p) mul(th[t], p, tbl[4096 + p * 16 + (t & 15)], true);
            radix16_dit_inv(th[t].re, th[t].im);
        }
        if (!exchange(th, kTile4096, s4096_a_write, s4096_a_read, true)) { printf("FAIL 4096 inverse exchange A\n"); return 1; }
    }
    for (int t = 0; t < n_threads; ++t) {
        for (int p = 0; p < 16; ++p) mul(th[t], p, tbl[p * n_threads + t], true);
        radix16_dit_inv(th[t].re, th[t].im); posttwist16_inv(th[t].re, th[t].im);
    }
    double werr = 0;
    for (int t = 0; t < n_threads; ++t)
        for (int m = 0; m < 16; ++m) {
            werr = std::max(werr, std::abs(th[t].re[m] - M * orig[t].re[m]));
            werr = std::max(werr, std::abs(th[t].im[m] - M * orig[t].im[m]));
        }
    if (werr > 1e-3) { printf("FAIL inverse(forward(x)) != M x (M = %d): %g\n", M, werr); return 1; }
    printf("M = %4d: forward max error %.3Lg, roundtrip max error %.3g\n", M, worst, werr);
    return 0;
}

int main() {
    if (!conflict_free(16, s256_write) || !conflict_free(16, s256_read)) { printf("FAIL bank conflicts (256)\n"); return 1; }
    if (!conflict_free(256, s4096_a_write) || !conflict_free(256, s4096_a_read) || !conflict_free(256, s4096_b_write) || !conflict_free(256, s4096_b_read)) {
        printf("FAIL bank conflicts (4096)\n"); return 1;
    }
    for (int t = 0; t < 256; ++t)
        for (int g = 0; g < 16; ++g)
            for (int a : {s4096_a_read(t, g), s4096_b_write(t, g), s4096_b_read(t, g)})
                if (a / 272 != (t >> 4)) { printf("FAIL region confinement\n"); return 1; }
    if (check(256, 16) || check(4096, 256)) return 1;
    printf("OK\n");
    return 0;
}
