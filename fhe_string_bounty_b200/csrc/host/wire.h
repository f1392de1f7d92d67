// wire.h -- tfhe-rs 0.5 serialized objects (bincode 1.3.3, fixed-width little-endian integers, u64 sequence lengths, u32 enum
// variant indices: what `bincode::serialize` and safe_serialize's DefaultOptions::with_fixint_encoding produce,
// tfhe/src/safe_deserialization.rs:13-34) for the two things that cross the wire on the KS-PBS path:
//
//   shortint::CompressedServerKey                         shortint/server_key/compressed.rs:10-17,43-55
//     key_switching_key: SeededLweKeyswitchKey<Vec<u64>>  core_crypto/entities/seeded_lwe_keyswitch_key.rs:10-21
//         data, decomp_base_log, decomp_level_count, output_lwe_size, compression_seed, ciphertext_modulus
//     bootstrapping_key: enum { Classic(SeededLweBootstrapKey) = 0, MultiBit { seeded_bsk, deterministic_execution } = 1 }
//         SeededLweBootstrapKey { ggsw_list }             entities/seeded_lwe_bootstrap_key.rs:15-24
//         SeededLweMultiBitBootstrapKey { ggsw_list, grouping_factor }      entities/seeded_lwe_multi_bit_bootstrap_key.rs:15-25
//         SeededGgswCiphertextList { data, glwe_size, polynomial_size, decomp_base_log, decomp_level_count, compression_seed,
//                                    ciphertext_modulus }                   entities/seeded_ggsw_ciphertext_list.rs:11-23
//     message_modulus, carry_modulus, max_degree (usize newtypes), ciphertext_modulus, pbs_order (enum, parameters.rs:233-245)
//     CompressionSeed { seed: u128 }                      commons/math/random/generator.rs:27-39
//     CiphertextModulus -> { modulus: u128 (0 = native), scalar_bits: usize }   commons/ciphertext_modulus.rs:41-64
//   shortint::Ciphertext { ct: LweCiphertext { data, ciphertext_modulus }, degree, noise_level, message_modulus, carry_modulus,
//                          pbs_order }                    shortint/ciphertext/mod.rs:261-270, entities/lwe_ciphertext.rs:500-507
//   a radix ciphertext is BaseRadixCiphertext { blocks: Vec<Ciphertext> }   integer/ciphertext/mod.rs:18-30
//
// No tfhe-rs build exists in this environment, so there is no real blob to pin the layout against: it is restated from the serde
// derives above (field order = declaration order) -- "parity unpinned", see DESIGN.md.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace tbw {

struct Reader {
    const uint8_t *p;
    size_t len, off = 0;
    Reader(const uint8_t *bytes, size_t n) : p(bytes), len(n) {}
    void need(size_t n) const {
        if (n > len - off) throw std::runtime_error("wire: truncated input at byte " + std::to_string(off));
    }
    uint8_t u8() { need(1); return p[off++]; }
    uint32_t u32() { need(4); uint32_t v; std::memcpy(&v, p + off, 4); off += 4; return v; }
    uint64_t u64() { need(8); uint64_t v; std::memcpy(&v, p + off, 8); off += 8; return v; }
    void u128(uint8_t out[16]) { need(16); std::memcpy(out, p + off, 16); off += 16; }
    // Vec<u64>: returns the byte offset of the first element, advances past it
    size_t vec_u64(uint64_t &count) {
        count = u64();
        if (count > (len - off) / 8) throw std::runtime_error("wire: sequence length exceeds the input");
        const size_t at = off;
        off += (size_t)count * 8;
        return at;
    }
    std::string str() {
        const uint64_t n = u64();
        need((size_t)n);
        std::string s(reinterpret_cast<const char *>(p + off), (size_t)n);
        off += (size_t)n;
        return s;
    }
    // CiphertextModulus<u64>: only the native modulus 2^64 is supported by the engine
    void native_modulus_u64(const char *what) {
        uint8_t m[16];
        u128(m);
        const uint64_t bits = u64();
        bool zero = true;
        for (int i = 0; i < 16; ++i) zero = zero && m[i] == 0;
        if (bits != 64) throw std::runtime_error(std::string("wire: ") + what + ": expected 64-bit scalars");
        if (!zero) throw std::runtime_error(std::string("wire: ") + what + ": only the native ciphertext modulus is supported");
    }
};

struct ServerKeyView {
    uint32_t lwe_dim = 0, glwe_dim = 0, poly_size = 0, pbs_base_log = 0, pbs_level = 0, ks_base_log = 0, ks_level = 0,
             grouping_factor = 0, msg_mod = 0, carry_mod = 0;
    uint32_t pbs_order = 0, deterministic = 0;
    uint64_t max_degree = 0;
    uint8_t ksk_seed[16] = {0}, bsk_seed[16] = {0};
    size_t ksk_off = 0, ksk_len = 0, bsk_off = 0, bsk_len = 0;   // byte offsets / word counts of the two body arrays
};

inline void read_seeded_ggsw_list(Reader &r, ServerKeyView &v) {
    uint64_t n = 0;
    v.bsk_off = r.vec_u64(n);
    v.bsk_len = (size_t)n;
    const uint64_t glwe_size = r.u64(), poly = r.u64(), base_log = r.u64(), level = r.u64();
    r.u128(v.bsk_seed);
    r.native_modulus_u64("bootstrap key");
    if (glwe_size < 2 || poly == 0 || level == 0 || glwe_size > 64 || poly > (1u << 20) || level > 64 || base_log > 64)
        throw std::runtime_error("wire: implausible bootstrap key dimensions");
    v.glwe_dim = (uint32_t)(glwe_size - 1);
    v.poly_size = (uint32_t)poly;
    v.pbs_base_log = (uint32_t)base_log;
    v.pbs_level = (uint32_t)level;
}

// bincode::serialize(&CompressedServerKey)
inline ServerKeyView parse_compressed_server_key(const uint8_t *bytes, size_t len) {
    Reader r(bytes, len);
    ServerKeyView v;
    uint64_t n = 0;
    v.ksk_off = r.vec_u64(n);
    v.ksk_len = (size_t)n;
    const uint64_t ks_base_log = r.u64(), ks_level = r.u64(), out_lwe_size = r.u64();
    r.u128(v.ksk_seed);
    r.native_modulus_u64("keyswitch key");
    if (ks_level == 0 || ks_level > 64 || ks_base_log > 64 || out_lwe_size < 2 || out_lwe_size > (1u << 20))
        throw std::runtime_error("wire: implausible keyswitch key dimensions");
    v.ks_base_log = (uint32_t)ks_base_log;
    v.ks_level = (uint32_t)ks_level;
    v.lwe_dim = (uint32_t)(out_lwe_size - 1);
    const uint32_t variant = r.u32();
    if (variant == 0) {
        read_seeded_ggsw_list(r, v);
    } else if (variant == 1) {
        read_seeded_ggsw_list(r, v);
        const uint64_t g = r.u64();
        if (g == 0 || g > 8) throw std::runtime_error("wire: implausible grouping factor");
        v.grouping_factor = (uint32_t)g;
        v.deterministic = r.u8() ? 1u : 0u;
    } else {
        throw std::runtime_error("wire: unknown ShortintCompressedBootstrappingKey variant");
    }
    v.msg_mod = (uint32_t)r.u64();
    v.carry_mod = (uint32_t)r.u64();
    v.max_degree = r.u64();
    r.native_modulus_u64("server key");
    v.pbs_order = r.u32();
    if (v.pbs_order > 1) throw std::runtime_error("wire: unknown PBSOrder variant");
    if (r.off != len) throw std::runtime_error("wire: trailing bytes after CompressedServerKey");
    // cross-checks between the containers and the scalar fields
    const size_t in_dim = (size_t)v.glwe_dim * v.poly_size;
    if (v.ksk_len != in_dim * v.ks_level) throw std::runtime_error("wire: keyswitch key body count does not match k*N*level");
    const size_t per_ggsw = (size_t)v.pbs_level * (v.glwe_dim + 1) * v.poly_size;
    const size_t n_ggsw = v.grouping_factor ? ((size_t)(v.lwe_dim / v.grouping_factor) << v.grouping_factor) : v.lwe_dim;
    if (v.grouping_factor && v.lwe_dim % v.grouping_factor) throw std::runtime_error("wire: lwe dimension not a multiple of the grouping factor");
    if (v.bsk_len != n_ggsw * per_ggsw) throw std::runtime_error("wire: bootstrap key body count does not match the dimensions");
    return v;
}

struct CiphertextMeta { uint64_t degree, noise_level, msg_mod, carry_mod; uint32_t pbs_order; };

inline void read_ciphertext(Reader &r, std::vector<uint64_t> &lwe, size_t &lwe_len, std::vector<CiphertextMeta> &meta) {
    uint64_t n = 0;
    const size_t at = r.vec_u64(n);
    if (lwe_len == 0) lwe_len = (size_t)n;
    if ((size_t)n != lwe_len || n == 0) throw std::runtime_error("wire: ciphertexts of different (or zero) LWE size");
    const size_t base = lwe.size();
    lwe.resize(base + (size_t)n);
    std::memcpy(lwe.data() + base, r.p + at, (size_t)n * 8);
    r.native_modulus_u64("ciphertext");
    CiphertextMeta m;
    m.degree = r.u64(); m.noise_level = r.u64(); m.msg_mod = r.u64(); m.carry_mod = r.u64(); m.pbs_order = r.u32();
    if (m.pbs_order > 1) throw std::runtime_error("wire: unknown PBSOrder variant");
    meta.push_back(m);
}

// is_radix = 0: one shortint::Ciphertext; 1: BaseRadixCiphertext { blocks: Vec<Ciphertext> }
inline void parse_ciphertexts(const uint8_t *bytes, size_t len, bool is_radix, std::vector<uint64_t> &lwe, size_t &lwe_len,
                              std::vector<CiphertextMeta> &meta) {
    Reader r(bytes, len);
    lwe.clear(); meta.clear(); lwe_len = 0;
    if (is_radix) {
        const uint64_t n = r.u64();
        if (n > len / 8) throw std::runtime_error("wire: block count exceeds the input");
        for (uint64_t i = 0; i < n; ++i) read_ciphertext(r, lwe, lwe_len, meta);
    } else {
        read_ciphertext(r, lwe, lwe_len, meta);
    }
    if (r.off != len) throw std::runtime_error("wire: trailing bytes after the ciphertext(s)");
}

struct Writer {
    std::vector<uint8_t> b;
    void u32(uint32_t v) { const uint8_t *q = reinterpret_cast<const uint8_t *>(&v); b.insert(b.end(), q, q + 4); }
    void u64(uint64_t v) { const uint8_t *q = reinterpret_cast<const uint8_t *>(&v); b.insert(b.end(), q, q + 8); }
    void native_modulus_u64() { for (int i = 0; i < 16; ++i) b.push_back(0); u64(64); }
};

inline std::vector<uint8_t> write_ciphertexts(const uint64_t *lwe, size_t lwe_len, const CiphertextMeta *meta, size_t n, bool is_radix) {
    Writer w;
    if (is_radix) w.u64(n);
    for (size_t i = 0; i < n; ++i) {
        w.u64(lwe_len);
        const uint8_t *q = reinterpret_cast<const uint8_t *>(lwe + i * lwe_len);
        w.b.insert(w.b.end(), q, q + lwe_len * 8);
        w.native_modulus_u64();
        w.u64(meta[i].degree); w.u64(meta[i].noise_level); w.u64(meta[i].msg_mod); w.u64(meta[i].carry_mod); w.u32(meta[i].pbs_order);
    }
    return w.b;
}

}  // namespace tbw
