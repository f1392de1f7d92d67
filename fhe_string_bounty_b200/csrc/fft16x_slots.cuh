// fft16x_slots.cuh -- exchange-slot arithmetic, frequency maps and twiddle tables of the 16-points-per-thread FFTs of pbs_n512.cu (256
// complex points: 16 threads, radix 16 x 16) and pbs_n8192.cu (4096 points: 256 threads, radix 16 x 16 x 16).  Host + device, so that
// tests/cpu_mirror/fft16x_mirror.cpp emulates both transforms thread by thread from the very functions the kernels use.
#pragma once
#include <cmath>

#include "fft16_core.cuh"

namespace tb16x {
using tb16::brev4;

// ---- 256 points (N = 512): point j = T + 16 m; exchange (thread T, register p) -> (thread p, register T); tile rows padded to 17 -------
constexpr int kTile256 = 16 * 17;
TB_HD constexpr int s256_write(int T, int p) { return p * 17 + T; }
TB_HD constexpr int s256_read(int T, int u) { return T * 17 + u; }
// frequency held by (thread T', register pv) after the forward transform
TB_HD constexpr int freq256(int Tp, int pv) { return brev4(Tp) + 16 * brev4(pv); }

// ---- 4096 points (N = 8192): point j = T + 256 m, T = u + 16 v; 16 regions of 272 slots, one per half-warp after exchange A ------------
constexpr int kTile4096 = 16 * 272;
TB_HD constexpr int s4096_a_write(int T, int p1) { return 272 * p1 + T; }                          // (T, p1) -> region p1
TB_HD constexpr int s4096_a_read(int Tp, int v) { return 272 * (Tp >> 4) + (Tp & 15) + 16 * v; }   // thread T' = u + 16 p1, register v
TB_HD constexpr int s4096_b_write(int Tp, int p2) { return 272 * (Tp >> 4) + 17 * p2 + (Tp & 15); }
TB_HD constexpr int s4096_b_read(int Tpp, int u) { return 272 * (Tpp >> 4) + 17 * (Tpp & 15) + u; } // thread T'' = p2 + 16 p1, register u
TB_HD constexpr int freq4096(int Tpp, int p3) { return brev4(Tpp >> 4) + 16 * brev4(Tpp & 15) + 256 * brev4(p3); }

}  // namespace tb16x

// T1[p * 16 + T] = exp(i pi T (1 - 4 brev4(p)) / 512): 256 complex values
static inline void tb16x_make_table_512(double *t) {
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int p = 0; p < 16; ++p)
        for (int T = 0; T < 16; ++T) {
            long e = ((long)T * (1 - 4 * (long)tb16::brev4(p))) % 1024;
            if (e < 0) e += 1024;
            t[2 * (p * 16 + T)] = (double)cosl(pi * (long double)e / 512.0L);
            t[2 * (p * 16 + T) + 1] = (double)sinl(pi * (long double)e / 512.0L);
        }
}
// T1[p1 * 256 + T] = exp(i pi T (1 - 4 brev4(p1)) / 8192) (4096 entries), then T2[p2 * 16 + u] = exp(-2 pi i u brev4(p2) / 256) (256 entries)
static inline void tb16x_make_table_8192(double *t) {
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int p = 0; p < 16; ++p)
        for (int T = 0; T < 256; ++T) {
            long e = ((long)T * (1 - 4 * (long)tb16::brev4(p))) % 16384;
            if (e < 0) e += 16384;
            t[2 * (p * 256 + T)] = (double)cosl(pi * (long double)e / 8192.0L);
            t[2 * (p * 256 + T) + 1] = (double)sinl(pi * (long double)e / 8192.0L);
        }
    for (int p = 0; p < 16; ++p)
        for (int u = 0; u < 16; ++u) {
            const int e = (u * tb16::brev4(p)) % 256;
            t[2 * (4096 + p * 16 + u)] = (double)cosl(-2.0L * pi * e / 256.0L);
            t[2 * (4096 + p * 16 + u) + 1] = (double)sinl(-2.0L * pi * e / 256.0L);
        }
}
