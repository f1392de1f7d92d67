"""Oracle restatement (CPU, test infrastructure only) of the integer-radix trees that feed the string ops, written
block-by-block the way the reference does it (one apply_lookup_table per block), plus a generic executor that runs a
program recorded by the product's host layer with the oracle's CPU KS-PBS.

Reference (tfhe/src/integer/server_key/): radix_parallel/comparison.rs:10-83, radix_parallel/scalar_comparison.rs:104-240,
comparator.rs:52-133,193-279,389-464,957-971,1103-1126."""
from __future__ import annotations

import numpy as np

from . import oracle as O

U64 = np.uint64


class ShortintServerKey:
    """shortint::ServerKey over the oracle primitives (server_key/mod.rs:383-476, bivariate_pbs.rs:71-182, add.rs:520-524,
    scalar_mul.rs:520-536, scalar_add.rs:211-218)."""

    def __init__(self, sk: O.ServerKey):
        self.sk = sk
        self.p = sk.p
        self.delta = (1 << 63) // (self.p.msg_mod * self.p.carry_mod)
        self.pbs_count = 0

    def lut(self, f):
        return self.sk.generate_lookup_table(f)[0]

    def lut_bivariate(self, f):
        return self.sk.generate_lookup_table_bivariate(f)[0]

    def apply_lookup_table(self, ct, acc):
        self.pbs_count += 1
        return self.sk.ks_pbs_batch(ct[None, :], acc[None, :])[0]

    def apply_lookup_table_many(self, cts, accs):
        """rayon par_iter stand-in: independent blocks, each with its own accumulator"""
        cts = np.stack(cts)
        uniq = {}
        idx = np.zeros(len(accs), dtype=np.uint32)
        table = []
        for i, a in enumerate(accs):
            k = a.tobytes()
            if k not in uniq:
                uniq[k] = len(table)
                table.append(a)
            idx[i] = uniq[k]
        self.pbs_count += len(cts)
        return list(self.sk.ks_pbs_batch(cts, np.stack(table), idx))

    @staticmethod
    def add(a, b):
        return a + b

    @staticmethod
    def sub(a, b):
        return a - b

    @staticmethod
    def scalar_mul(a, k):
        return a * U64(k)

    def scalar_add(self, a, v):
        r = a.copy()
        with np.errstate(over="ignore"):
            r[-1] += U64((v * self.delta) % 2**64)
        return r

    def create_trivial(self, v):
        r = np.zeros(self.p.big_dim + 1, dtype=U64)
        r[-1] = U64((v % self.p.msg_mod) * self.delta)
        return r


class IntegerServerKey:
    IS_INFERIOR, IS_EQUAL, IS_SUPERIOR = 0, 1, 2

    def __init__(self, key: ShortintServerKey):
        self.key = key
        self.p = key.p

    # scalar_comparison.rs:147-198
    def are_all_comparisons_block_true(self, blocks):
        if not blocks:
            return self.key.create_trivial(1)
        max_value = self.p.msg_mod * self.p.carry_mod - 1
        while len(blocks) > 1:
            sums, accs = [], []
            for i in range(0, len(blocks), max_value):
                chunk = blocks[i:i + max_value]
                s = chunk[0].copy()
                for o in chunk[1:]:
                    s = self.key.add(s, o)
                sums.append(s)
                accs.append(self.key.lut(lambda x, n=len(chunk): int(x == n)))
            blocks = self.key.apply_lookup_table_many(sums, accs)
        return blocks[0]

    # scalar_comparison.rs:200-240
    def is_at_least_one_comparisons_block_true(self, blocks):
        if not blocks:
            return self.key.create_trivial(1)
        max_value = self.p.msg_mod * self.p.carry_mod - 1
        acc = self.key.lut(lambda x: int(x != 0))
        while len(blocks) > 1:
            sums = []
            for i in range(0, len(blocks), max_value):
                chunk = blocks[i:i + max_value]
                s = chunk[0].copy()
                for o in chunk[1:]:
                    s = self.key.add(s, o)
                sums.append(s)
            blocks = self.key.apply_lookup_table_many(sums, [acc] * len(sums))
        return blocks[0]

    # comparison.rs:10-33 / 35-83
    def unchecked_eq(self, lhs, rhs):
        m = self.p.msg_mod
        acc = self.key.lut_bivariate(lambda x, y: int(x == y))
        packed = [self.key.add(self.key.scalar_mul(l, m), r) for l, r in zip(lhs, rhs)]
        return self.are_all_comparisons_block_true(self.key.apply_lookup_table_many(packed, [acc] * len(packed)))

    def unchecked_ne(self, lhs, rhs):
        m = self.p.msg_mod
        acc = self.key.lut_bivariate(lambda x, y: int(x != y))
        packed = [self.key.add(self.key.scalar_mul(l, m), r) for l, r in zip(lhs, rhs)]
        if not packed:
            return self.key.create_trivial(0)
        return self.is_at_least_one_comparisons_block_true(self.key.apply_lookup_table_many(packed, [acc] * len(packed)))

    # scalar_comparison.rs:104-139
    def pack_block_chunk(self, chunk):
        if len(chunk) == 1:
            return chunk[0].copy()
        return self.key.add(self.key.scalar_mul(chunk[1], self.p.msg_mod), chunk[0])

    # comparator.rs:389-464 (carry_modulus >= message_modulus branch) + 193-221 + 257-279
    def unchecked_compare(self, lhs, rhs):
        sign_lut = self.key.lut(lambda x: int(x != 0))
        diffs = []
        for i in range(0, len(lhs), 2):
            pl = self.pack_block_chunk(lhs[i:i + 2])
            pr = self.pack_block_chunk(rhs[i:i + 2])
            diffs.append(self.key.sub(pl, pr))
        signs = [self.key.scalar_add(s, 1) for s in self.key.apply_lookup_table_many(diffs, [sign_lut] * len(diffs))]
        table = [0, 0, 0, 0, 0, 1, 2, 2, 2, 2, 2]
        red = self.key.lut(lambda x: table[x] if x < 11 else 0)
        while len(signs) != 1:
            packed = [self.key.add(self.key.scalar_mul(signs[i + 1], 4), signs[i]) for i in range(0, len(signs) - 1, 2)]
            nxt = self.key.apply_lookup_table_many(packed, [red] * len(packed))
            if len(signs) % 2 == 1:
                nxt.append(signs[-1])
            signs = nxt
        return signs[0]

    # comparator.rs:957-971, 1103-1126
    def map_sign(self, sign, h):
        return self.key.apply_lookup_table(sign, self.key.lut(lambda x: int(h(x))))

    def unchecked_lt(self, a, b):
        return self.map_sign(self.unchecked_compare(a, b), lambda x: x == self.IS_INFERIOR)

    def unchecked_le(self, a, b):
        return self.map_sign(self.unchecked_compare(a, b), lambda x: x in (self.IS_INFERIOR, self.IS_EQUAL))

    def unchecked_gt(self, a, b):
        return self.map_sign(self.unchecked_compare(a, b), lambda x: x == self.IS_SUPERIOR)

    def unchecked_ge(self, a, b):
        return self.map_sign(self.unchecked_compare(a, b), lambda x: x in (self.IS_SUPERIOR, self.IS_EQUAL))


def run_program(ir, sk: O.ServerKey, inputs: np.ndarray) -> np.ndarray:
    """Execute a program recorded by the product's host layer (fhe_string_bounty_b200.host.ProgramIR) with the ORACLE's
    leveled arithmetic and CPU KS-PBS.  The accumulators are rebuilt here with the oracle's generate_lookup_table."""
    p = sk.p
    L = p.big_dim + 1
    arena = np.zeros((ir.n_slots, L), dtype=U64)
    inputs = np.ascontiguousarray(inputs, dtype=U64).reshape(-1, L)
    assert inputs.shape[0] == ir.n_inputs
    arena[: ir.n_inputs] = inputs
    accs = np.stack([sk.generate_lookup_table(lambda x, t=t: int(t[x]))[0] for t in ir.lut_tables]) if len(ir.lut_tables) else None
    for lv in range(len(ir.level_lin_off) - 1):
        for li in range(ir.level_lin_off[lv], ir.level_lin_off[lv + 1]):
            out, tb, te = (int(v) for v in ir.lin[li])
            acc = np.zeros(L, dtype=U64)
            for t in range(tb, te):
                acc += arena[ir.term_slot[t]] * U64(int(ir.term_coef[t]) % 2**64)
            with np.errstate(over="ignore"):
                acc[-1] += ir.lin_body[li]
            arena[out] = acc
        p0, p1 = int(ir.level_pbs_off[lv]), int(ir.level_pbs_off[lv + 1])
        if p1 > p0:
            jobs = ir.pbs[p0:p1]
            res = sk.ks_pbs_batch(arena[jobs[:, 0]], accs, jobs[:, 2].astype(np.uint32))
            arena[jobs[:, 1]] = res
    return arena[ir.outputs]


def encrypt_radix(ck: O.ClientKey, value: int, n_blocks: int) -> list[np.ndarray]:
    """integer/encryption.rs:46-83: little-endian msg_mod-ary digits"""
    m = ck.p.msg_mod
    out = []
    for _ in range(n_blocks):
        out.append(ck.encrypt(value % m))
        value //= m
    return out


def blocks_per_char(msg_mod: int) -> int:
    """a char is an FheUint8: ceil(8 / log2(message_modulus)) blocks (integer/encryption.rs:69-83); 4 for the 2-bit sets"""
    bits = int(msg_mod).bit_length() - 1
    assert bits >= 1 and (1 << bits) == msg_mod
    return -(-8 // bits)


def encrypt_string(ck: O.ClientKey, s: bytes) -> np.ndarray:
    """little-endian blocks of log2(message_modulus) bits per char: 4 2-bit blocks for the MESSAGE_2 sets
    (examples/regex_engine/ciphertext.rs:19-22)"""
    blocks = []
    n = blocks_per_char(ck.p.msg_mod)
    for ch in s:
        blocks.extend(encrypt_radix(ck, ch, n))
    return np.stack(blocks) if blocks else np.zeros((0, ck.p.big_dim + 1), dtype=U64)


def decrypt_radix(ck: O.ClientKey, blocks) -> int:
    m = ck.p.msg_mod
    v = 0
    for i, b in enumerate(blocks):
        v += ck.decrypt(b) * m**i
    return v


def decrypt_string(ck: O.ClientKey, blocks) -> bytes:
    n = blocks_per_char(ck.p.msg_mod)
    return bytes(decrypt_radix(ck, blocks[i:i + n]) & 0xFF for i in range(0, len(blocks), n))
