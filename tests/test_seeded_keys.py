"""Seeded (compressed) server keys, SURVEY.md section 8(f) N2: the CPU restatement of concrete-csprng's AES-128 CTR stream and of tfhe's
seeded_*_decompression.rs, pinned by the known answers the reference's own tests hold."""
import ctypes as C

import numpy as np

# concrete-csprng/src/generators/implem/aesni/block_cipher.rs:188-205 (FIPS-197 appendix A.1 / B)
CIPHER_KEY = "000102030405060708090a0b0c0d0e0f"
KEY_SCHEDULE = [
    "000102030405060708090a0b0c0d0e0f", "d6aa74fdd2af72fadaa678f1d6ab76fe", "b692cf0b643dbdf1be9bc5006830b3fe",
    "b6ff744ed2c2c9bf6c590cbf0469bf41", "47f7f7bc95353e03f96c32bcfd058dfd", "3caaa3e8a99f9deb50f3af57adf622aa",
    "5e390f7df7a69296a7553dc10aa31f6b", "14f9701ae35fe28c440adf4d4ea9c026", "47438735a41c65b9e016baf4aebf7ad2",
    "549932d1f08557681093ed9cbe2c974e", "13111d7fe3944a17f307a78b4d2b30c5",
]
PLAINTEXT = "00112233445566778899aabbccddeeff"
CIPHERTEXT = "69c4e0d86a7b0430d8cdb78070b4c55a"


def _b(h):
    return np.frombuffer(bytes.fromhex(h), dtype=np.uint8).copy()


def test_aes128_reference_known_answers(orc):
    """test_generate_key_schedule / test_encrypt_many_messages of the reference (aesni/block_cipher.rs:207-229, soft/block_cipher.rs:84-113)."""
    L = orc.lib()
    rk = np.zeros(176, dtype=np.uint8)
    L.orc_aes128_expand_key(_b(CIPHER_KEY), rk)
    assert [rk[16 * i:16 * i + 16].tobytes().hex() for i in range(11)] == KEY_SCHEDULE
    ct = np.zeros(16, dtype=np.uint8)
    L.orc_aes128_encrypt_block(rk, _b(PLAINTEXT), ct)
    assert ct.tobytes().hex() == CIPHERTEXT


def test_csprng_stream_layout(orc):
    """generic.rs:27-37: a fresh generator starts at TableIndex::SECOND (aes 0, byte 1); states.rs/index.rs: table byte 16*A + b is byte b of
    AES_seed(A as little-endian u128); generic.rs:79-118 + prop_fork: forks hand out consecutive ranges, so skipping == forking."""
    L = orc.lib()
    seed = orc.seed_bytes(0x0F0E0D0C0B0A09080706050403020100)      # to_ne_bytes -> 00 01 02 ... 0f = the FIPS key
    assert seed.tobytes().hex() == CIPHER_KEY
    rk = np.zeros(176, dtype=np.uint8)
    L.orc_aes128_expand_key(seed, rk)
    blocks = []
    for a in range(4):
        ctr = np.frombuffer(int(a).to_bytes(16, "little"), dtype=np.uint8).copy()
        out = np.zeros(16, dtype=np.uint8)
        L.orc_aes128_encrypt_block(rk, ctr, out)
        blocks.append(out)
    table = np.concatenate(blocks)
    got = np.zeros(40, dtype=np.uint8)
    L.orc_csprng_generate_bytes(seed, 0, got, 40)
    assert np.array_equal(got, table[1:41])                          # first output byte is table byte 1
    # "children concatenated == parent" (prop_fork): a child that starts after 3 children of 7 bytes sees the parent's bytes 21..
    child = np.zeros(7, dtype=np.uint8)
    L.orc_csprng_generate_bytes(seed, 3 * 7, child, 7)
    assert np.array_equal(child, got[21:28])
    # uniform.rs:13-24: u64::from_le_bytes of 8 consecutive bytes
    words = np.zeros(4, dtype=np.uint64)
    L.orc_csprng_mask_words(seed, 1, words, 4)
    assert [int(w) for w in words] == [int.from_bytes(got[8 * (j + 1):8 * (j + 2)].tobytes(), "little") for j in range(4)]


def test_seeded_server_key_roundtrip_toy(orc, toy_keys):
    """compress (test-side key generation with the seeded mask stream) -> decompress (the reference's algorithm): masks are the stream, bodies
    are copied, and the decompressed keys work: keyswitch + PBS decrypt to the LUT values for every message."""
    p, ck, sk = toy_keys
    L = orc.lib()
    csk = orc.CompressedServerKey(ck, sk, ksk_seed=0x1234567890ABCDEF1122334455667788, bsk_seed=0x0FEDCBA987654321)
    assert csk.ksk_bodies.size == p.glwe_dim * p.poly_size * p.ks_level
    assert csk.bsk_bodies.size == p.lwe_dim * p.pbs_level * (p.glwe_dim + 1) * p.poly_size
    ksk, bsk = csk.decompress()
    # layout: LWE row c of the KSK = [n stream words | body c]; GLWE row g of the BSK = [k*N stream words | body polynomial g]
    n, N, k = p.lwe_dim, p.poly_size, p.glwe_dim
    rows = ksk.reshape(-1, n + 1)
    want = np.zeros(3 * n, dtype=np.uint64)
    L.orc_csprng_mask_words(csk.ksk_seed, 0, want, 3 * n)
    assert np.array_equal(rows[:3, :n].ravel(), want) and np.array_equal(rows[:, n], csk.ksk_bodies)
    g = bsk.reshape(-1, (k + 1) * N)
    want = np.zeros(k * N, dtype=np.uint64)
    L.orc_csprng_mask_words(csk.bsk_seed, 5 * k * N, want, k * N)
    assert np.array_equal(g[5, :k * N], want) and np.array_equal(g[:, k * N:].ravel(), csk.bsk_bodies)
    assert not np.array_equal(ksk, sk.ksk) and not np.array_equal(bsk, sk.bsk)
    # the decompressed key is a valid server key
    sk2 = orc.ServerKey.__new__(orc.ServerKey)
    sk2.p, sk2.ksk, sk2.bsk, sk2._fourier = p, ksk, bsk, None
    acc, _ = sk2.generate_lookup_table(lambda x: (3 * x + 1) % (p.msg_mod * p.carry_mod))
    for m in range(p.msg_mod * p.carry_mod):
        ct = ck.encrypt_with_carry(m)
        out = sk2.pbs(sk2.keyswitch(ct), acc)
        assert ck.decrypt_message_and_carry(out) == (3 * m + 1) % (p.msg_mod * p.carry_mod)
