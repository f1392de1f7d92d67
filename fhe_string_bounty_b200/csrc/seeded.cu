// seeded.cu -- device-side decompression of seeded (compressed) keys.
//
// A tfhe-rs CompressedServerKey (shortint/server_key/mod.rs:935-1023) holds, per key, one 128-bit compression seed and the ciphertext
// BODIES only; the masks are re-drawn from concrete-csprng's AES-128 CTR stream (core_crypto/algorithms/
// seeded_lwe_bootstrap_key_decompression.rs, seeded_ggsw_ciphertext_list_decompression.rs:9-49, seeded_glwe_ciphertext_decompression.rs:6-45,
// seeded_lwe_keyswitch_key_decompression.rs:6-27, seeded_lwe_ciphertext_list_decompression.rs:9-60).  Every fork in that call tree hands
// consecutive byte ranges to its children (concrete-csprng generators/aes_ctr/generic.rs:79-118), a fresh generator starts at table byte 1
// (generic.rs:27-37, TableIndex::SECOND) and a u64 is from_le_bytes of 8 consecutive bytes (commons/math/random/uniform.rs:13-24), so
//     mask word j of the key (every mask coefficient, in storage order) = LE u64 of table bytes [1 + 8j, 9 + 8j),
//     table byte 16*A + b = AES-128_seed(A as 16 little-endian bytes)[b]      (implem/soft/block_cipher.rs:14-36)
// which is embarrassingly parallel: one thread per AES block, 257 blocks per CTA staged in shared memory, 512 mask words out.
// Halves the host->device bytes of a key upload (109 MB -> 55 MB for the 2_2 parameters) and removes the host-side decompression.
#include "kernels.h"
#include <cstdint>
#include <cstring>

namespace tbs {

// FIPS-197 S-box
#define TB_SBOX_BYTES \
    0x63,0x7c,0x77,0x7b,0xf2,0x6b,0x6f,0xc5,0x30,0x01,0x67,0x2b,0xfe,0xd7,0xab,0x76,0xca,0x82,0xc9,0x7d,0xfa,0x59,0x47,0xf0, \
    0xad,0xd4,0xa2,0xaf,0x9c,0xa4,0x72,0xc0,0xb7,0xfd,0x93,0x26,0x36,0x3f,0xf7,0xcc,0x34,0xa5,0xe5,0xf1,0x71,0xd8,0x31,0x15, \
    0x04,0xc7,0x23,0xc3,0x18,0x96,0x05,0x9a,0x07,0x12,0x80,0xe2,0xeb,0x27,0xb2,0x75,0x09,0x83,0x2c,0x1a,0x1b,0x6e,0x5a,0xa0, \
    0x52,0x3b,0xd6,0xb3,0x29,0xe3,0x2f,0x84,0x53,0xd1,0x00,0xed,0x20,0xfc,0xb1,0x5b,0x6a,0xcb,0xbe,0x39,0x4a,0x4c,0x58,0xcf, \
    0xd0,0xef,0xaa,0xfb,0x43,0x4d,0x33,0x85,0x45,0xf9,0x02,0x7f,0x50,0x3c,0x9f,0xa8,0x51,0xa3,0x40,0x8f,0x92,0x9d,0x38,0xf5, \
    0xbc,0xb6,0xda,0x21,0x10,0xff,0xf3,0xd2,0xcd,0x0c,0x13,0xec,0x5f,0x97,0x44,0x17,0xc4,0xa7,0x7e,0x3d,0x64,0x5d,0x19,0x73, \
    0x60,0x81,0x4f,0xdc,0x22,0x2a,0x90,0x88,0x46,0xee,0xb8,0x14,0xde,0x5e,0x0b,0xdb,0xe0,0x32,0x3a,0x0a,0x49,0x06,0x24,0x5c, \
    0xc2,0xd3,0xac,0x62,0x91,0x95,0xe4,0x79,0xe7,0xc8,0x37,0x6d,0x8d,0xd5,0x4e,0xa9,0x6c,0x56,0xf4,0xea,0x65,0x7a,0xae,0x08, \
    0xba,0x78,0x25,0x2e,0x1c,0xa6,0xb4,0xc6,0xe8,0xdd,0x74,0x1f,0x4b,0xbd,0x8b,0x8a,0x70,0x3e,0xb5,0x66,0x48,0x03,0xf6,0x0e, \
    0x61,0x35,0x57,0xb9,0x86,0xc1,0x1d,0x9e,0xe1,0xf8,0x98,0x11,0x69,0xd9,0x8e,0x94,0x9b,0x1e,0x87,0xe9,0xce,0x55,0x28,0xdf, \
    0x8c,0xa1,0x89,0x0d,0xbf,0xe6,0x42,0x68,0x41,0x99,0x2d,0x0f,0xb0,0x54,0xbb,0x16

static const uint8_t h_sbox[256] = {TB_SBOX_BYTES};
__constant__ uint8_t d_sbox[256] = {TB_SBOX_BYTES};

struct RoundKeys { uint32_t w[44]; };   // 11 round keys, column c of round r = w[4r + c] (byte row i at bits 8i)

// FIPS-197 5.2 on the host (176 bytes; the seed is the cipher key, soft/block_cipher.rs:16)
static RoundKeys expand_key(const uint8_t key[16]) {
    uint8_t rk[176];
    std::memcpy(rk, key, 16);
    uint8_t rcon = 1;
    for (int i = 16; i < 176; i += 4) {
        uint8_t t[4] = {rk[i - 4], rk[i - 3], rk[i - 2], rk[i - 1]};
        if (i % 16 == 0) {
            const uint8_t t0 = t[0];
            t[0] = h_sbox[t[1]] ^ rcon; t[1] = h_sbox[t[2]]; t[2] = h_sbox[t[3]]; t[3] = h_sbox[t0];
            rcon = (uint8_t)((rcon << 1) ^ ((rcon >> 7) * 0x1b));
        }
        for (int b = 0; b < 4; ++b) rk[i + b] = rk[i - 16 + b] ^ t[b];
    }
    RoundKeys out;
    for (int i = 0; i < 44; ++i) out.w[i] = (uint32_t)rk[4 * i] | ((uint32_t)rk[4 * i + 1] << 8) | ((uint32_t)rk[4 * i + 2] << 16) | ((uint32_t)rk[4 * i + 3] << 24);
    return out;
}

__device__ __forceinline__ uint32_t rotr8(uint32_t x, int bytes) { return __funnelshift_r(x, x, 8 * bytes); }

// AES-128 of the 128-bit little-endian counter `a` (high 64 bits zero); state column c = bytes in[4c .. 4c+3]
__device__ __forceinline__ uint4 aes128_ctr_block(const uint8_t *sbox, const RoundKeys &rk, uint64_t a) {
    uint32_t s0 = (uint32_t)a ^ rk.w[0], s1 = (uint32_t)(a >> 32) ^ rk.w[1], s2 = rk.w[2], s3 = rk.w[3];
#pragma unroll
    for (int round = 1; round <= 10; ++round) {
        // SubBytes + ShiftRows: new column c, row r = S[old column (c + r) & 3, row r]
        const uint32_t t0 = sbox[s0 & 255] | (sbox[(s1 >> 8) & 255] << 8) | (sbox[(s2 >> 16) & 255] << 16) | (sbox[s3 >> 24] << 24);
        const uint32_t t1 = sbox[s1 & 255] | (sbox[(s2 >> 8) & 255] << 8) | (sbox[(s3 >> 16) & 255] << 16) | (sbox[s0 >> 24] << 24);
        const uint32_t t2 = sbox[s2 & 255] | (sbox[(s3 >> 8) & 255] << 8) | (sbox[(s0 >> 16) & 255] << 16) | (sbox[s1 >> 24] << 24);
        const uint32_t t3 = sbox[s3 & 255] | (sbox[(s0 >> 8) & 255] << 8) | (sbox[(s1 >> 16) & 255] << 16) | (sbox[s2 >> 24] << 24);
        if (round < 10) {
            // MixColumns, four bytes at a time: b = 2a ^ rot8(2a ^ a) ^ rot16(a) ^ rot24(a)
            auto mix = [](uint32_t c) {
                const uint32_t x = ((c & 0x7f7f7f7fu) << 1) ^ (((c >> 7) & 0x01010101u) * 0x1bu);
                return x ^ rotr8(x ^ c, 1) ^ rotr8(c, 2) ^ rotr8(c, 3);
            };
            s0 = mix(t0) ^ rk.w[4 * round];
            s1 = mix(t1) ^ rk.w[4 * round + 1];
            s2 = mix(t2) ^ rk.w[4 * round + 2];
            s3 = mix(t3) ^ rk.w[4 * round + 3];
        } else {
            s0 = t0 ^ rk.w[40]; s1 = t1 ^ rk.w[41]; s2 = t2 ^ rk.w[42]; s3 = t3 ^ rk.w[43];
        }
    }
    return make_uint4(s0, s1, s2, s3);
}

constexpr int BLOCKS_PER_CTA = 256;                    // AES blocks per CTA (+1 for the one-byte skew)
constexpr int WORDS_PER_CTA = 2 * BLOCKS_PER_CTA;      // mask words per CTA

// dst row g = [mask_len stream words | body_len words copied from bodies]
__global__ void __launch_bounds__(BLOCKS_PER_CTA)
seeded_expand_kernel(RoundKeys rk, uint64_t *__restrict__ dst, const uint64_t *__restrict__ bodies, unsigned long long n_rows,
                     unsigned int mask_len, unsigned int body_len) {
    __shared__ uint8_t sbox[256];
    __shared__ __align__(16) uint64_t tab[2 * (BLOCKS_PER_CTA + 1)];
    const int t = threadIdx.x;
    sbox[t] = d_sbox[t];
    __syncthreads();
    const unsigned long long a0 = (unsigned long long)blockIdx.x * BLOCKS_PER_CTA;
    reinterpret_cast<uint4 *>(tab)[t] = aes128_ctr_block(sbox, rk, a0 + t);
    if (t == 0) reinterpret_cast<uint4 *>(tab)[BLOCKS_PER_CTA] = aes128_ctr_block(sbox, rk, a0 + BLOCKS_PER_CTA);
    __syncthreads();
    const unsigned long long total = n_rows * mask_len;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = t + h * BLOCKS_PER_CTA;                                   // local word: table bytes [1 + 8j, 9 + 8j)
        const unsigned long long e = (unsigned long long)blockIdx.x * WORDS_PER_CTA + j;
        if (e < total) {
            const uint64_t v = (tab[j] >> 8) | (tab[j + 1] << 56);
            const unsigned long long g = e / mask_len;
            dst[g * (mask_len + body_len) + (e - g * mask_len)] = v;
        }
    }
    // bodies: this CTA copies the body words with the same global indices as its mask words (grid-stride over the rest)
    const unsigned long long n_body = n_rows * body_len;
    for (unsigned long long e = (unsigned long long)blockIdx.x * BLOCKS_PER_CTA + t; e < n_body; e += (unsigned long long)gridDim.x * BLOCKS_PER_CTA) {
        const unsigned long long g = e / body_len;
        dst[g * (mask_len + body_len) + mask_len + (e - g * body_len)] = __ldg(bodies + e);
    }
}

}  // namespace tbs

namespace tbk {

cudaError_t launch_seeded_expand(const uint8_t seed[16], uint64_t *dst, const uint64_t *bodies, size_t n_rows, uint32_t mask_len,
                                 uint32_t body_len, cudaStream_t stream) {
    if (n_rows == 0) return cudaSuccess;
    const tbs::RoundKeys rk = tbs::expand_key(seed);
    const size_t total = n_rows * mask_len;
    const unsigned grid = (unsigned)((total + tbs::WORDS_PER_CTA - 1) / tbs::WORDS_PER_CTA);
    tbs::seeded_expand_kernel<<<grid, tbs::BLOCKS_PER_CTA, 0, stream>>>(rk, dst, bodies, (unsigned long long)n_rows, mask_len, body_len);
    return cudaGetLastError();
}

}  // namespace tbk
