// Lane-by-lane CPU emulation of the warp-level FFT in fhe_string_bounty_b200/csrc/fft_core.cuh.
// Checks (1) the forward transform against the direct definition Z_k = sum_j z_j w^j W^(jk) and the
// (thread, register) -> frequency map, (2) inverse(forward(x)) == 1024 * x, (3) the transpose indexing.
// Built and run by tests/test_cpu_mirror.py (g++ -O1 -ffp-contract=off).
#include "../../fhe_string_bounty_b200/csrc/fft_core.cuh"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace tb;
typedef std::complex<long double> cld;

struct Warp {
    double re[32][32], im[32][32];  // [lane][register]
};

static void transpose(Warp &w) {
    static double xb[kXposeWords];
    for (int part = 0; part < 2; ++part) {
        for (int l = 0; l < 32; ++l)
            for (int r = 0; r < 32; ++r) xb[xpose_write_idx(l, r)] = part ? w.im[l][r] : w.re[l][r];
        for (int l = 0; l < 32; ++l)
            for (int r = 0; r < 32; ++r) (part ? w.im[l][r] : w.re[l][r]) = xb[xpose_read_idx(l, r)];
    }
}

int main() {
    std::vector<cplx> tbl(1024);
    tb_make_twiddle_table(reinterpret_cast<double *>(tbl.data()));
    auto load = [&](int i) { return tbl[i]; };
    const long double pi = 3.14159265358979323846264338327950288L;

    // transpose really is (lane, reg) <-> (reg, lane)
    {
        Warp w;
        for (int l = 0; l < 32; ++l) for (int r = 0; r < 32; ++r) { w.re[l][r] = l * 100 + r; w.im[l][r] = -(l * 100 + r); }
        transpose(w);
        for (int l = 0; l < 32; ++l) for (int r = 0; r < 32; ++r)
            if (w.re[l][r] != r * 100 + l || w.im[l][r] != -(r * 100 + l)) { printf("FAIL transpose\n"); return 1; }
    }

    srand(7);
    std::vector<cld> z(kM);
    Warp w;
    for (int j = 0; j < kM; ++j) {
        double a = (double)((rand() % (1 << 23)) - (1 << 22)), b = (double)((rand() % (1 << 23)) - (1 << 22));
        z[j] = cld(a, b);
        w.re[j & 31][j >> 5] = a;   // lane l = j mod 32, register m = j / 32
        w.im[j & 31][j >> 5] = b;
    }
    Warp orig = w;

    // forward
    for (int l = 0; l < 32; ++l) { pretwist_fwd(w.re[l], w.im[l]); radix32_dif(w.re[l], w.im[l]); twiddle_fwd(w.re[l], w.im[l], load, l); }
    transpose(w);
    for (int l = 0; l < 32; ++l) radix32_dif(w.re[l], w.im[l]);

    // direct definition (long double)
    long double max_err = 0, max_mag = 0;
    for (int t = 0; t < 32; ++t)
        for (int p = 0; p < 32; ++p) {
            const int k = freq_of(t, p);
            cld acc = 0;
            for (int j = 0; j < kM; ++j) {
                long e = ((long)j * (1 - 4 * (long)k)) % 4096;
                acc += z[j] * std::polar(1.0L, pi * (long double)e / 2048.0L);
            }
            long double err = std::abs(acc - cld(w.re[t][p], w.im[t][p]));
            if (err > max_err) max_err = err;
            if (std::abs(acc) > max_mag) max_mag = std::abs(acc);
        }
    printf("forward: max |err| = %.3Le (max |Z| = %.3Le, rel %.3Le)\n", max_err, max_mag, max_err / max_mag);
    if (max_err / max_mag > 1e-14L) { printf("FAIL forward\n"); return 1; }

    // inverse
    for (int l = 0; l < 32; ++l) radix32_dit_inv(w.re[l], w.im[l]);
    transpose(w);
    for (int l = 0; l < 32; ++l) { twiddle_inv(w.re[l], w.im[l], load, l); radix32_dit_inv(w.re[l], w.im[l]); posttwist_inv(w.re[l], w.im[l]); }
    long double max_rt = 0;
    for (int l = 0; l < 32; ++l)
        for (int m = 0; m < 32; ++m) {
            long double er = fabsl((long double)w.re[l][m] / 1024.0L - orig.re[l][m]);
            long double ei = fabsl((long double)w.im[l][m] / 1024.0L - orig.im[l][m]);
            if (er > max_rt) max_rt = er;
            if (ei > max_rt) max_rt = ei;
        }
    printf("roundtrip: max |err| = %.3Le on inputs of magnitude 2^22\n", max_rt);
    if (max_rt > 1e-8L) { printf("FAIL roundtrip\n"); return 1; }

    // integer helpers vs straightforward restatements
    for (int it = 0; it < 200000; ++it) {
        uint64_t x = ((uint64_t)rand() << 42) ^ ((uint64_t)rand() << 21) ^ (uint64_t)rand();
        if (it < 4) x = it == 0 ? 0 : it == 1 ? ~0ULL : it == 2 ? (1ULL << 40) : (1ULL << 63);
        for (int bl : {21, 23}) {
            // decomposer.rs:98-118 then iter.rs:120-127, level 1
            uint64_t shift = 64 - bl - 1, r = x >> shift; r += 1; r &= ~1ULL; r <<= shift;
            uint64_t state = r >> (64 - bl), mask = (1ULL << bl) - 1;
            uint64_t res = state & mask; state >>= bl;
            uint64_t carry = ((res - 1) | state) & res; carry >>= bl - 1;
            int64_t want = (int64_t)(res - (carry << bl));
            if (want != (int64_t)signed_digit_l1(x, bl)) { printf("FAIL digit %llx bl=%d: %lld vs %d\n", (unsigned long long)x, bl, (long long)want, signed_digit_l1(x, bl)); return 1; }
        }
        uint64_t ms = x >> (64 - 11 - 2); ms += 1; ms >>= 1;
        if (ms != modulus_switch_2n(x)) { printf("FAIL modswitch\n"); return 1; }
    }
    // rotation source map vs definition of poly * X^a
    for (uint32_t a : {0u, 1u, 5u, 2047u, 2048u, 2049u, 4095u, 3000u}) {
        std::vector<long> poly(kN), out(kN, 0);
        for (int j = 0; j < kN; ++j) poly[j] = j + 1;
        for (int j = 0; j < kN; ++j) {   // X^j * X^a
            uint32_t d = (j + a) % (2 * kN);
            if (d < (uint32_t)kN) out[d] += poly[j]; else out[d - kN] -= poly[j];
        }
        for (int j = 0; j < kN; ++j) {
            int src; bool neg; rot_src(j, a, src, neg);
            if ((neg ? -poly[src] : poly[src]) != out[j]) { printf("FAIL rot a=%u j=%d\n", a, j); return 1; }
        }
    }
    printf("OK\n");
    return 0;
}
