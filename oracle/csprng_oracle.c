/* csprng_oracle.c -- TEST INFRASTRUCTURE ONLY (see tfhe_oracle.h): CPU restatement of the reference's seeded-key path.
 *
 *   concrete-csprng (vendored in the reference tree, /concrete-csprng):
 *     src/generators/implem/soft/block_cipher.rs:14-36   AES-128, key = seed.to_ne_bytes(), plaintext = aes_index.to_ne_bytes()
 *     src/generators/aes_ctr/generic.rs:27-37,79-118     the generator starts at TableIndex::SECOND = (aes 0, byte 1); a fork hands
 *                                                        consecutive byte ranges to its children (prop_fork: children concatenated
 *                                                        == parent stream), so nested forks never reorder anything
 *     src/generators/aes_ctr/states.rs:20-45, index.rs   byte (aes_index A, byte_index b) of the table = AES_k(A)[b]
 *   tfhe core_crypto:
 *     commons/math/random/uniform.rs:13-24               u64 = from_le_bytes of 8 consecutive stream bytes
 *     commons/generators/encryption/mask_random_generator.rs:64-127,347-400   byte counts of the nested forks (all exact)
 *     algorithms/seeded_lwe_bootstrap_key_decompression.rs, seeded_ggsw_ciphertext_list_decompression.rs:9-49,
 *     seeded_ggsw_ciphertext_decompression.rs:8-55, seeded_glwe_ciphertext_decompression.rs:6-45   masks from the stream, bodies copied
 *     algorithms/seeded_lwe_keyswitch_key_decompression.rs:6-27, seeded_lwe_ciphertext_list_decompression.rs:9-60
 *   => mask word j (counting every mask coefficient of the key in storage order) = LE u64 of stream bytes [1 + 8j, 9 + 8j).
 *
 * Pinned by the reference's own known answers: FIPS-197 key schedule and ciphertext
 * (implem/aesni/block_cipher.rs:188-229, implem/soft/block_cipher.rs:84-113) -- tests/test_oracle_kat.py.
 */
#include "tfhe_oracle.h"
#include <stdlib.h>
#include <string.h>

static const uint8_t SBOX[256] = {
    0x63,0x7c,0x77,0x7b,0xf2,0x6b,0x6f,0xc5,0x30,0x01,0x67,0x2b,0xfe,0xd7,0xab,0x76,0xca,0x82,0xc9,0x7d,0xfa,0x59,0x47,0xf0,
    0xad,0xd4,0xa2,0xaf,0x9c,0xa4,0x72,0xc0,0xb7,0xfd,0x93,0x26,0x36,0x3f,0xf7,0xcc,0x34,0xa5,0xe5,0xf1,0x71,0xd8,0x31,0x15,
    0x04,0xc7,0x23,0xc3,0x18,0x96,0x05,0x9a,0x07,0x12,0x80,0xe2,0xeb,0x27,0xb2,0x75,0x09,0x83,0x2c,0x1a,0x1b,0x6e,0x5a,0xa0,
    0x52,0x3b,0xd6,0xb3,0x29,0xe3,0x2f,0x84,0x53,0xd1,0x00,0xed,0x20,0xfc,0xb1,0x5b,0x6a,0xcb,0xbe,0x39,0x4a,0x4c,0x58,0xcf,
    0xd0,0xef,0xaa,0xfb,0x43,0x4d,0x33,0x85,0x45,0xf9,0x02,0x7f,0x50,0x3c,0x9f,0xa8,0x51,0xa3,0x40,0x8f,0x92,0x9d,0x38,0xf5,
    0xbc,0xb6,0xda,0x21,0x10,0xff,0xf3,0xd2,0xcd,0x0c,0x13,0xec,0x5f,0x97,0x44,0x17,0xc4,0xa7,0x7e,0x3d,0x64,0x5d,0x19,0x73,
    0x60,0x81,0x4f,0xdc,0x22,0x2a,0x90,0x88,0x46,0xee,0xb8,0x14,0xde,0x5e,0x0b,0xdb,0xe0,0x32,0x3a,0x0a,0x49,0x06,0x24,0x5c,
    0xc2,0xd3,0xac,0x62,0x91,0x95,0xe4,0x79,0xe7,0xc8,0x37,0x6d,0x8d,0xd5,0x4e,0xa9,0x6c,0x56,0xf4,0xea,0x65,0x7a,0xae,0x08,
    0xba,0x78,0x25,0x2e,0x1c,0xa6,0xb4,0xc6,0xe8,0xdd,0x74,0x1f,0x4b,0xbd,0x8b,0x8a,0x70,0x3e,0xb5,0x66,0x48,0x03,0xf6,0x0e,
    0x61,0x35,0x57,0xb9,0x86,0xc1,0x1d,0x9e,0xe1,0xf8,0x98,0x11,0x69,0xd9,0x8e,0x94,0x9b,0x1e,0x87,0xe9,0xce,0x55,0x28,0xdf,
    0x8c,0xa1,0x89,0x0d,0xbf,0xe6,0x42,0x68,0x41,0x99,0x2d,0x0f,0xb0,0x54,0xbb,0x16};

const uint8_t *orc_aes_sbox(void) { return SBOX; }

static inline uint8_t xtime(uint8_t x) { return (uint8_t)((x << 1) ^ ((x >> 7) * 0x1b)); }

/* FIPS-197 5.2: 11 round keys of 16 bytes */
void orc_aes128_expand_key(const uint8_t key[16], uint8_t rk[176]) {
    memcpy(rk, key, 16);
    uint8_t rcon = 1;
    for (int i = 16; i < 176; i += 4) {
        uint8_t t[4] = {rk[i - 4], rk[i - 3], rk[i - 2], rk[i - 1]};
        if (i % 16 == 0) {
            const uint8_t t0 = t[0];
            t[0] = SBOX[t[1]] ^ rcon; t[1] = SBOX[t[2]]; t[2] = SBOX[t[3]]; t[3] = SBOX[t0];
            rcon = xtime(rcon);
        }
        for (int b = 0; b < 4; b++) rk[i + b] = rk[i - 16 + b] ^ t[b];
    }
}

/* FIPS-197 5.1 (state byte s[r + 4c] = in[r + 4c]) */
void orc_aes128_encrypt_block(const uint8_t rk[176], const uint8_t in[16], uint8_t out[16]) {
    uint8_t s[16], t[16];
    for (int i = 0; i < 16; i++) s[i] = in[i] ^ rk[i];
    for (int round = 1; round <= 10; round++) {
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) t[r + 4 * c] = SBOX[s[r + 4 * ((c + r) & 3)]];   /* SubBytes + ShiftRows */
        if (round < 10) {
            for (int c = 0; c < 4; c++) {
                const uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                s[4 * c + 0] = xtime(a0) ^ (xtime(a1) ^ a1) ^ a2 ^ a3;
                s[4 * c + 1] = a0 ^ xtime(a1) ^ (xtime(a2) ^ a2) ^ a3;
                s[4 * c + 2] = a0 ^ a1 ^ xtime(a2) ^ (xtime(a3) ^ a3);
                s[4 * c + 3] = (xtime(a0) ^ a0) ^ a1 ^ a2 ^ xtime(a3);
            }
        } else {
            memcpy(s, t, 16);
        }
        for (int i = 0; i < 16; i++) s[i] ^= rk[16 * round + i];
    }
    memcpy(out, s, 16);
}

/* bytes [first, first + n) of the AES-CTR table of `seed` (table byte 16*A + b = AES_seed(A as 16 LE bytes)[b]) */
void orc_csprng_table_bytes(const uint8_t seed[16], uint64_t first, uint8_t *out, size_t n) {
    uint8_t rk[176], ctr[16], blk[16];
    orc_aes128_expand_key(seed, rk);
    uint64_t have = (uint64_t)-1;
    for (size_t i = 0; i < n; i++) {
        const uint64_t pos = first + i, a = pos >> 4;
        if (a != have) {
            memset(ctr, 0, 16);
            for (int b = 0; b < 8; b++) ctr[b] = (uint8_t)(a >> (8 * b));
            orc_aes128_encrypt_block(rk, ctr, blk);
            have = a;
        }
        out[i] = blk[pos & 15];
    }
}

/* what a freshly seeded generator hands out (it starts at table byte 1), skipping `skip` bytes first */
void orc_csprng_generate_bytes(const uint8_t seed[16], uint64_t skip, uint8_t *out, size_t n) {
    orc_csprng_table_bytes(seed, 1 + skip, out, n);
}

/* mask words [first_word, first_word + n): uniform.rs from_le_bytes over the generator's byte stream */
void orc_csprng_mask_words(const uint8_t seed[16], uint64_t first_word, uint64_t *out, size_t n) {
    uint8_t *buf = (uint8_t *)malloc(n ? 8 * n : 1);
    orc_csprng_generate_bytes(seed, 8 * first_word, buf, 8 * n);
    for (size_t j = 0; j < n; j++) {
        uint64_t v = 0;
        for (int b = 0; b < 8; b++) v |= (uint64_t)buf[8 * j + b] << (8 * b);
        out[j] = v;
    }
    free(buf);
}

/* number of ciphertexts (GLWE rows / LWE rows) of the two keys, and their mask / body lengths */
static size_t bsk_rows(const orc_params *p) {
    const size_t n_ggsw = p->grouping_factor ? (size_t)(p->lwe_dim / p->grouping_factor) << p->grouping_factor : p->lwe_dim;
    return n_ggsw * p->pbs_level * (p->glwe_dim + 1);
}
size_t orc_seeded_bsk_len(const orc_params *p) { return bsk_rows(p) * p->poly_size; }
size_t orc_seeded_ksk_len(const orc_params *p) { return (size_t)p->glwe_dim * p->poly_size * p->ks_level; }

/* seeded_ggsw_ciphertext_list_decompression.rs:9-49 -> seeded_ggsw_ciphertext_decompression.rs:8-55 ->
 * seeded_glwe_ciphertext_decompression.rs:6-45: GLWE row g = [k mask polynomials from the stream | stored body polynomial] */
void orc_decompress_seeded_bsk(const orc_params *p, const uint8_t seed[16], const uint64_t *bodies, uint64_t *bsk_std) {
    const size_t N = p->poly_size, k = p->glwe_dim, rows = bsk_rows(p);
    #pragma omp parallel for schedule(static)
    for (long g = 0; g < (long)rows; g++) {
        uint64_t *glwe = bsk_std + (size_t)g * (k + 1) * N;
        orc_csprng_mask_words(seed, (uint64_t)g * k * N, glwe, k * N);
        memcpy(glwe + k * N, bodies + (size_t)g * N, N * sizeof(uint64_t));
    }
}

/* seeded_lwe_keyswitch_key_decompression.rs:6-27 -> seeded_lwe_ciphertext_list_decompression.rs:9-60: LWE c = [n mask words | body] */
void orc_decompress_seeded_ksk(const orc_params *p, const uint8_t seed[16], const uint64_t *bodies, uint64_t *ksk) {
    const size_t n = p->lwe_dim, cts = orc_seeded_ksk_len(p);
    #pragma omp parallel for schedule(static)
    for (long c = 0; c < (long)cts; c++) {
        uint64_t *lwe = ksk + (size_t)c * (n + 1);
        orc_csprng_mask_words(seed, (uint64_t)c * n, lwe, n);
        lwe[n] = bodies[c];
    }
}

/* Test-side key generation: turn a standard key into the seeded key a client would have produced with compression seed `seed`
 * (same plaintexts, same noise, masks replaced by the seeded stream): body' = body - <mask, s> + <mask', s>. */
static void negacyclic_mul_binary_acc(uint64_t *body, const uint64_t *mask, const uint64_t *key, size_t N, int sign) {
    for (size_t t = 0; t < N; t++) {
        if (!key[t]) continue;
        if (sign > 0) {
            for (size_t j = 0; j < N - t; j++) body[j + t] += mask[j];
            for (size_t j = N - t; j < N; j++) body[j + t - N] -= mask[j];
        } else {
            for (size_t j = 0; j < N - t; j++) body[j + t] -= mask[j];
            for (size_t j = N - t; j < N; j++) body[j + t - N] += mask[j];
        }
    }
}
void orc_compress_bsk(const orc_params *p, const uint64_t *glwe_sk, const uint8_t seed[16], const uint64_t *bsk_std, uint64_t *bodies) {
    const size_t N = p->poly_size, k = p->glwe_dim, rows = bsk_rows(p);
    #pragma omp parallel for schedule(dynamic, 8)
    for (long g = 0; g < (long)rows; g++) {
        const uint64_t *glwe = bsk_std + (size_t)g * (k + 1) * N;
        uint64_t *body = bodies + (size_t)g * N;
        uint64_t *mask = (uint64_t *)malloc(k * N * sizeof(uint64_t));
        memcpy(body, glwe + k * N, N * sizeof(uint64_t));
        orc_csprng_mask_words(seed, (uint64_t)g * k * N, mask, k * N);
        for (size_t i = 0; i < k; i++) {
            negacyclic_mul_binary_acc(body, glwe + i * N, glwe_sk + i * N, N, -1);
            negacyclic_mul_binary_acc(body, mask + i * N, glwe_sk + i * N, N, +1);
        }
        free(mask);
    }
}
void orc_compress_ksk(const orc_params *p, const uint64_t *small_sk, const uint8_t seed[16], const uint64_t *ksk, uint64_t *bodies) {
    const size_t n = p->lwe_dim, cts = orc_seeded_ksk_len(p);
    #pragma omp parallel for schedule(static)
    for (long c = 0; c < (long)cts; c++) {
        const uint64_t *lwe = ksk + (size_t)c * (n + 1);
        uint64_t *mask = (uint64_t *)malloc(n * sizeof(uint64_t));
        orc_csprng_mask_words(seed, (uint64_t)c * n, mask, n);
        uint64_t b = lwe[n];
        for (size_t i = 0; i < n; i++) b += (mask[i] - lwe[i]) * small_sk[i];
        bodies[c] = b;
        free(mask);
    }
}
