// host/padded.h -- the part of the FheString surface whose RESULT LENGTH IS SECRET: strings in the null-padded model.
//
// A padded string has a public capacity and a secret length: the content is followed by zero bytes up to the capacity (once a zero
// byte appears every later byte is zero).  That is the model the bounty's FheString uses to hide lengths ("API
// surface of the missing string module"); the reference snapshot itself only has the building blocks -- FheUint8 chars with
// to_upper / to_lower in docs/tutorials/ascii_fhe_string.md:81-154, and a zero-terminated-free Vec<RadixCiphertext> in
// examples/regex_engine/ciphertext.rs:4-22 / execution.rs:63-215 -- so, as for host/strings.h, only DECRYPTED results are comparable
// (against Rust's str semantics: len, is_empty, ==, <, trim_start / trim_end / trim with char::is_whitespace on ASCII, strip_prefix,
// strip_suffix, +, repeat, contains / starts_with / ends_with).
//
// Everything is composed from the radix methods of host/radix.h (count trees, comparator, cmux, carry-free one-hot sums) and recorded
// into the same level-batched programs: wide levels of independent KS-PBS jobs, which is what the engine is built for.  Secret shifts
// (trim_start, concat, strip_prefix with a padded pattern) are barrel shifters: one cmux level per bit of the shift amount
// (radix_parallel/cmux.rs:211-250), the amount itself being the radix index of a one-hot boundary flag, never a sum that could overflow
// a block.
#pragma once
#include "strings.h"

namespace tbh {

class PaddedStringServerKey {
  public:
    explicit PaddedStringServerKey(Program &prog) : pg(prog), isk(prog), ssk(prog), p(prog.params()) {
        ssk.require_two_bit_blocks("null-padded strings");
    }

    // ---- chars -----------------------------------------------------------------------------------------------------------------
    Radix trivial_char(unsigned char ch) {
        Radix c;
        for (size_t b = 0; b < 4; ++b) c.push_back(pg.create_trivial((ch >> (2 * b)) & 3));
        return c;
    }
    FheString extend(const FheString &s, size_t cap) {
        FheString r = s;
        while (r.len() < cap) r.chars.push_back(trivial_char(0));
        return r;
    }
    // b0 + b1 + b2 + b3 (<= 12): zero iff the char is the padding byte
    Ct block_sum(const Radix &c) {
        Ct s = c[0];
        for (size_t b = 1; b < c.size(); ++b) s = pg.unchecked_add(s, c[b]);
        return s;
    }
    BooleanBlock nonzero(const Radix &c) { return pg.pbs(block_sum(c), [](uint64_t x) { return uint64_t(x != 0); }); }
    BooleanBlock is_zero(const Radix &c) { return pg.pbs(block_sum(c), [](uint64_t x) { return uint64_t(x == 0); }); }
    // keep ? c : 0   (zero_out_if, radix_parallel/cmux.rs:281-)
    Radix keep_if(const Radix &c, const Ct &keep) {
        Radix r;
        for (auto &blk : c) r.push_back(pg.pbs_bivariate(blk, keep, [](uint64_t v, uint64_t k) { return (k & 1) ? v : uint64_t(0); }));
        return r;
    }
    // cond ? a : b per block (cmux.rs:211-250: two zero_out_if, add, message_extract)
    Radix select(const Ct &cond, const Radix &a, const Radix &b) { return isk.if_then_else(cond, a, b); }

    // ---- scans over boolean flags ---------------------------------------------------------------------------------------------
    // out[i] = OR_{j < i} f[j] (exclusive) or OR_{j <= i} f[j] (inclusive) in THREE levels: flags are cut into blocks of 14,
    //   level 1   any_k      = [sum of block k != 0]
    //   level 2   earlier_k  = [some block before k matched]     (running sums of <= 15 cleaned flags; a cleaned carry every group)
    //   level 3   out[i]     = [(flags of its block up to i) + earlier_k != 0]      (<= 14 + 1)
    std::vector<Ct> prefix_or(const std::vector<Ct> &f, bool inclusive) {
        const size_t n = f.size();
        std::vector<Ct> out(n);
        if (n == 0) return out;
        constexpr size_t BLK = 14;
        const size_t n_blk = (n + BLK - 1) / BLK, max_sum = p.total_mod() - 1;
        std::vector<Ct> any(n_blk), earlier(n_blk);
        for (size_t k = 0; k < n_blk; ++k) {
            Ct sum = f[k * BLK];
            for (size_t i = k * BLK + 1; i < std::min(n, (k + 1) * BLK); ++i) sum = pg.unchecked_add(sum, f[i]);
            any[k] = n_blk > 1 ? pg.pbs(sum, [](uint64_t x) { return uint64_t(x != 0); }) : sum;
        }
        earlier[0] = pg.create_trivial(0);
        Ct carry = pg.create_trivial(0);
        for (size_t g0 = 0; g0 < n_blk;) {
            const size_t g1 = std::min(n_blk, g0 + (g0 == 0 ? max_sum : max_sum - 1));
            Ct run = carry;
            for (size_t k = g0; k < g1; ++k) {
                if (k > 0) earlier[k] = pg.pbs(run, [](uint64_t x) { return uint64_t(x != 0); });
                run = pg.unchecked_add(run, any[k]);
            }
            if (g1 < n_blk) carry = pg.pbs(run, [](uint64_t x) { return uint64_t(x != 0); });
            g0 = g1;
        }
        for (size_t i = 0; i < n; ++i) {
            const size_t k = i / BLK;
            Ct y = earlier[k];
            for (size_t j = k * BLK; j < i + (inclusive ? 1 : 0); ++j) y = pg.unchecked_add(y, f[j]);
            out[i] = y.is_trivial() ? y : pg.pbs(y, [](uint64_t x) { return uint64_t(x != 0); });
        }
        return out;
    }
    // radix digits of the position of the single 1 in `onehot` (position w counts as value_of(w)); at most one flag is set, so the
    // selected digits are summed without carries, 15 at a time with a cleaning PBS (the chunking of scalar_comparison.rs:155-170)
    Radix onehot_to_radix(const std::vector<Ct> &onehot, size_t n_digits, const std::function<uint64_t(size_t)> &value_of) {
        const size_t max_value = p.total_mod() - 1;
        Radix digits;
        for (size_t d = 0; d < n_digits; ++d) {
            std::vector<Ct> terms;
            for (size_t w = 0; w < onehot.size(); ++w) {
                const uint64_t digit = (value_of(w) >> (2 * d)) & 3;
                if (digit == 0) continue;
                terms.push_back(pg.pbs(onehot[w], [digit](uint64_t x) { return (x & 1) ? digit : uint64_t(0); }));
            }
            if (terms.empty()) { digits.push_back(pg.create_trivial(0)); continue; }
            while (terms.size() > 1) {
                std::vector<Ct> next;
                for (size_t i = 0; i < terms.size(); i += max_value) {
                    const size_t len = std::min(max_value, terms.size() - i);
                    Ct sum = terms[i];
                    for (size_t j = 1; j < len; ++j) sum = pg.unchecked_add(sum, terms[i + j]);
                    sum.degree = p.msg_mod - 1;   // at most one term is non-zero
                    next.push_back(pg.pbs(sum, [this](uint64_t x) { return x % p.msg_mod; }));
                }
                terms.swap(next);
            }
            digits.push_back(terms[0]);
        }
        return digits;
    }
    static size_t digits_for(size_t max_value) {
        size_t d = 1;
        while ((size_t(1) << (2 * d)) <= max_value) ++d;
        return d;
    }
    // binary bits of a radix number, little-endian (2 per digit)
    std::vector<Ct> bits_of(const Radix &digits, size_t n_bits) {
        std::vector<Ct> bits;
        for (size_t d = 0; d < digits.size() && bits.size() < n_bits; ++d) {
            bits.push_back(pg.pbs(digits[d], [](uint64_t x) { return x & 1; }));
            if (bits.size() < n_bits) bits.push_back(pg.pbs(digits[d], [](uint64_t x) { return (x >> 1) & 1; }));
        }
        return bits;
    }
    // barrel shifter: s shifted by the secret amount sum_t bit_t 2^t (left: towards index 0, dropping chars; right: towards the end,
    // zero-filling), one cmux level per bit
    FheString shift(const FheString &s, const std::vector<Ct> &bits, bool left) {
        FheString cur = s;
        const size_t n = s.len();
        for (size_t t = 0; t < bits.size(); ++t) {
            const size_t k = size_t(1) << t;
            FheString nxt = cur;
            for (size_t i = 0; i < n; ++i) {
                const bool has_src = left ? (i + k < n) : (i >= k);
                if (has_src) nxt.chars[i] = select(bits[t], cur.chars[left ? i + k : i - k], cur.chars[i]);
                else nxt.chars[i] = keep_if(cur.chars[i], isk.boolean_bitnot(bits[t]));     // the shifted-in char is the padding byte
            }
            cur = nxt;
        }
        return cur;
    }

    // ---- len / is_empty ----------------------------------------------------------------------------------------------------------
    std::vector<Ct> nonzero_flags(const FheString &s) {
        std::vector<Ct> nz;
        for (auto &c : s.chars) nz.push_back(nonzero(c));
        return nz;
    }
    // boundary one-hot over positions 0 .. n: 1 at i == len
    std::vector<Ct> length_onehot(const std::vector<Ct> &nz) {
        const size_t n = nz.size();
        std::vector<Ct> b(n + 1);
        if (n == 0) { b[0] = pg.create_trivial(1); return b; }
        b[0] = isk.boolean_bitnot(nz[0]);
        for (size_t i = 1; i < n; ++i) {
            Ct y = pg.unchecked_add(nz[i - 1], pg.unchecked_scalar_mul(nz[i], 2));   // (prev, cur) = (1, 0) <=> y == 1
            b[i] = pg.pbs(y, [](uint64_t x) { return uint64_t(x == 1); });
        }
        b[n] = nz[n - 1];
        return b;
    }
    Radix len(const FheString &s) {
        return onehot_to_radix(length_onehot(nonzero_flags(s)), digits_for(s.len()), [](size_t w) { return uint64_t(w); });
    }
    BooleanBlock is_empty(const FheString &s) { return s.len() == 0 ? pg.create_trivial(1) : is_zero(s.chars[0]); }

    // ---- comparisons: the padding byte sorts below every char, so the unpadded circuits apply to the common capacity -------------------
    BooleanBlock eq(const FheString &a, const FheString &b) { const size_t n = std::max(a.len(), b.len()); return ssk.eq(extend(a, n), extend(b, n)); }
    BooleanBlock ne(const FheString &a, const FheString &b) { const size_t n = std::max(a.len(), b.len()); return ssk.ne(extend(a, n), extend(b, n)); }
    BooleanBlock lt(const FheString &a, const FheString &b) { const size_t n = std::max(a.len(), b.len()); return ssk.lt(extend(a, n), extend(b, n)); }
    BooleanBlock le(const FheString &a, const FheString &b) { const size_t n = std::max(a.len(), b.len()); return ssk.le(extend(a, n), extend(b, n)); }
    BooleanBlock gt(const FheString &a, const FheString &b) { const size_t n = std::max(a.len(), b.len()); return ssk.gt(extend(a, n), extend(b, n)); }
    BooleanBlock ge(const FheString &a, const FheString &b) { const size_t n = std::max(a.len(), b.len()); return ssk.ge(extend(a, n), extend(b, n)); }

    // ---- whitespace classes: 0 = other char, 1 = ASCII whitespace (U+0009 ..= U+000D, U+0020: char::is_whitespace), 2 = padding ------
    Ct char_class(const Radix &c) {
        Ct hi = pg.unchecked_add(pg.unchecked_scalar_mul(c[3], p.msg_mod), c[2]);
        Ct lo = pg.unchecked_add(pg.unchecked_scalar_mul(c[1], p.msg_mod), c[0]);
        Ct k = pg.pbs(hi, [](uint64_t x) { return x == 0 ? uint64_t(1) : (x == 2 ? uint64_t(2) : uint64_t(0)); });
        Ct r = pg.pbs(lo, [](uint64_t x) { return x == 0 ? uint64_t(2) : ((x >= 9 && x <= 13) ? uint64_t(1) : uint64_t(0)); });
        return pg.pbs_bivariate(k, r, [](uint64_t kk, uint64_t rr) {
            if (kk == 1 && rr == 2) return uint64_t(2);                        // 0x00
            if ((kk == 1 && rr == 1) || (kk == 2 && rr == 2)) return uint64_t(1);   // 0x09..0x0D, 0x20
            return uint64_t(0);
        });
    }
    FheString trim_end(const FheString &s) {
        const size_t n = s.len();
        if (n == 0) return s;
        std::vector<Ct> other(n);
        for (size_t i = 0; i < n; ++i) other[n - 1 - i] = pg.pbs(char_class(s.chars[i]), [](uint64_t x) { return uint64_t(x == 0); });
        std::vector<Ct> keep = prefix_or(other, true);      // over the reversed string: some non-whitespace char at or after i
        FheString out = s;
        for (size_t i = 0; i < n; ++i) out.chars[i] = keep_if(s.chars[i], keep[n - 1 - i]);
        return out;
    }
    FheString trim_start(const FheString &s) {
        const size_t n = s.len();
        if (n == 0) return s;
        std::vector<Ct> stop(n);                            // not whitespace: the first such char is where the result starts
        for (size_t i = 0; i < n; ++i) stop[i] = pg.pbs(char_class(s.chars[i]), [](uint64_t x) { return uint64_t(x != 1); });
        std::vector<Ct> before = prefix_or(stop, false);
        std::vector<Ct> first(n + 1);
        for (size_t i = 0; i < n; ++i) {
            Ct y = pg.unchecked_add(stop[i], pg.unchecked_scalar_mul(before[i], 2));
            first[i] = pg.pbs(y, [](uint64_t x) { return uint64_t(x == 1); });
        }
        {   // whitespace up to the capacity: everything goes
            Ct all = pg.unchecked_add(stop[n - 1], before[n - 1]);
            first[n] = pg.pbs(all, [](uint64_t x) { return uint64_t(x == 0); });
        }
        const size_t nd = digits_for(n);
        Radix amount = onehot_to_radix(first, nd, [](size_t w) { return uint64_t(w); });
        size_t n_bits = 1;
        while ((size_t(1) << n_bits) <= n) ++n_bits;
        return shift(s, bits_of(amount, n_bits), true);
    }
    FheString trim(const FheString &s) { return trim_end(trim_start(s)); }

    // ---- strip_prefix / strip_suffix with a clear pattern: (matched, string) -------------------------------------------------------------
    std::pair<BooleanBlock, FheString> strip_prefix(const FheString &s, const std::string &pat) {
        const size_t n = s.len(), m = pat.size();
        if (m == 0) return {pg.create_trivial(1), s};
        if (m > n) return {pg.create_trivial(0), s};
        BooleanBlock flag = ssk.starts_with(s, ssk.trivial_string(pat));
        FheString out = s;
        for (size_t i = 0; i < n; ++i)
            out.chars[i] = i + m < n ? select(flag, s.chars[i + m], s.chars[i]) : keep_if(s.chars[i], isk.boolean_bitnot(flag));
        return {flag, out};
    }
    // ends_with for a padded haystack: some window equals the pattern AND the string ends right after it
    BooleanBlock ends_with_clear(const FheString &s, const std::string &pat, const std::vector<Ct> &nz) {
        const size_t n = s.len(), m = pat.size();
        if (m == 0) return pg.create_trivial(1);
        if (m > n) return pg.create_trivial(0);
        FheString tp = ssk.trivial_string(pat);
        std::vector<Ct> per_window;
        for (size_t w = 0; w + m <= n; ++w) {
            std::vector<Ct> flags = isk.block_equalities(StringServerKey::concat(s, w, m), StringServerKey::concat(tp, 0, m));
            if (w + m < n) flags.push_back(isk.boolean_bitnot(nz[w + m]));
            per_window.push_back(isk.are_all_comparisons_block_true(flags));
        }
        return isk.is_at_least_one_comparisons_block_true(per_window);
    }
    std::pair<BooleanBlock, FheString> strip_suffix(const FheString &s, const std::string &pat) {
        const size_t n = s.len(), m = pat.size();
        if (m == 0) return {pg.create_trivial(1), s};
        if (m > n) return {pg.create_trivial(0), s};
        std::vector<Ct> nz = nonzero_flags(s);
        BooleanBlock flag = ends_with_clear(s, pat, nz);
        FheString out = s;
        for (size_t i = 0; i < n; ++i) {
            // char i belongs to the suffix iff the string ends within m chars of it: nz[i + m] == 0 (beyond the capacity: always)
            Ct keep = i + m < n ? pg.pbs_bivariate(flag, nz[i + m], [](uint64_t f, uint64_t z) { return uint64_t(!((f & 1) && !(z & 1))); })
                                : isk.boolean_bitnot(flag);
            out.chars[i] = keep_if(s.chars[i], keep);
        }
        return {flag, out};
    }

    // ---- pattern matching with a PADDED (secret-length) pattern ---------------------------------------------------------------------------
    // window w matches iff for every j: pat[j] is padding OR hay[w + j] == pat[j]  (chars beyond the haystack's capacity are padding)
    BooleanBlock window_match(const FheString &hay, const FheString &pat, size_t w, const std::vector<Ct> &pat_zero) {
        std::vector<Ct> ok;
        for (size_t j = 0; j < pat.len(); ++j) {
            if (w + j >= hay.len()) { ok.push_back(pat_zero[j]); continue; }
            std::vector<Ct> eqs = isk.block_equalities(hay.chars[w + j], pat.chars[j]);
            Ct y = pg.unchecked_scalar_mul(pat_zero[j], 4);
            for (auto &e : eqs) y = pg.unchecked_add(y, e);                 // 4 [pat[j] == 0] + #equal blocks  in [0, 8]
            ok.push_back(pg.pbs(y, [](uint64_t x) { return uint64_t(x >= 4); }));
        }
        return isk.are_all_comparisons_block_true(ok);
    }
    std::vector<Ct> zero_flags(const FheString &s) {
        std::vector<Ct> z;
        for (auto &c : s.chars) z.push_back(is_zero(c));
        return z;
    }
    BooleanBlock starts_with(const FheString &hay, const FheString &pat) {
        if (pat.len() == 0) return pg.create_trivial(1);
        return window_match(hay, pat, 0, zero_flags(pat));
    }
    BooleanBlock contains(const FheString &hay, const FheString &pat) {
        if (pat.len() == 0) return pg.create_trivial(1);
        std::vector<Ct> pz = zero_flags(pat), m;
        for (size_t w = 0; w < std::max<size_t>(hay.len(), 1); ++w) m.push_back(window_match(hay, pat, w, pz));
        return isk.is_at_least_one_comparisons_block_true(m);
    }
    // find / rfind with a padded pattern (str::find / str::rfind: byte index of the first / last match; the empty pattern matches at 0 and,
    // for rfind, at len).  Window w = 0 .. capacity: a match must start inside the string or right at its end (w == 0 or hay[w - 1] != 0),
    // which only matters for the empty pattern -- a non-empty one cannot match in the padding.  Index: radix digits of w, as string_find.
    std::pair<BooleanBlock, Radix> find(const FheString &hay, const FheString &pat, bool last) {
        const size_t n = hay.len();
        size_t idx_blocks = 1;
        while ((size_t(1) << (2 * idx_blocks)) < n + 1) ++idx_blocks;
        std::vector<Ct> pz = zero_flags(pat), m;
        for (size_t w = 0; w <= n; ++w) {
            Ct mw = w < n || pat.len() == 0 ? (pat.len() == 0 ? pg.create_trivial(1) : window_match(hay, pat, w, pz))
                                            : isk.are_all_comparisons_block_true(pz);            // w == n: only the empty pattern fits
            if (w > 0) mw = isk.boolean_bitand(mw, nonzero(hay.chars[w - 1]));
            m.push_back(mw);
        }
        if (last) std::reverse(m.begin(), m.end());
        return ssk.first_true(m, [=](size_t w) { return last ? n - w : w; }, idx_blocks);
    }

    // some suffix of the haystack equals the pattern as padded strings (the empty suffix included)
    BooleanBlock ends_with(const FheString &hay, const FheString &pat) {
        const size_t n = hay.len(), m = pat.len();
        std::vector<Ct> pz = zero_flags(pat), hz = zero_flags(hay), per;
        for (size_t w = 0; w <= n; ++w) {
            std::vector<Ct> flags;
            const size_t span = std::max(n - w, m);
            for (size_t j = 0; j < span; ++j) {
                const bool in_h = w + j < n, in_p = j < m;
                if (in_h && in_p) {
                    std::vector<Ct> eqs = isk.block_equalities(hay.chars[w + j], pat.chars[j]);
                    flags.insert(flags.end(), eqs.begin(), eqs.end());
                } else if (in_h) flags.push_back(hz[w + j]);
                else flags.push_back(pz[j]);
            }
            per.push_back(isk.are_all_comparisons_block_true(flags));
        }
        return isk.is_at_least_one_comparisons_block_true(per);
    }

    // ---- concat / repeat: the second operand moves right by the secret length of the first; the supports are disjoint, so the result is a
    // leveled sum of blocks -----------------------------------------------------------------------------------------------------------------------
    FheString concat(const FheString &a, const FheString &b) {
        const size_t na = a.len(), nb = b.len(), n = na + nb;
        if (na == 0) return b;
        if (nb == 0) return a;
        Radix la = len(a);
        size_t n_bits = 1;
        while ((size_t(1) << n_bits) <= na) ++n_bits;
        FheString bs = shift(extend(b, n), bits_of(la, n_bits), false);
        FheString ae = extend(a, n), out;
        for (size_t i = 0; i < n; ++i) {
            Radix c;
            for (size_t k = 0; k < 4; ++k) {
                Ct v = pg.unchecked_add(ae.chars[i][k], bs.chars[i][k]);
                v.degree = p.msg_mod - 1;        // at most one of the two blocks is non-zero
                c.push_back(v);
            }
            out.chars.push_back(c);
        }
        return out;
    }
    FheString repeat(const FheString &s, size_t count) {
        if (count == 0) { FheString e; return e; }
        FheString acc = s;
        for (size_t k = 1; k < count; ++k) acc = concat(acc, s);
        return acc;
    }

    Program &pg;
    IntegerServerKey isk;
    StringServerKey ssk;
    Params p;
};

}  // namespace tbh
