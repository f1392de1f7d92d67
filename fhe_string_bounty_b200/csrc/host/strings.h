// host/strings.h -- encrypted ASCII string operations composed from the radix methods (host/radix.h).
//
// The reference snapshot has no FheString module: only the tutorial type
// FheAsciiString{bytes: Vec<FheUint8>} with to_upper/to_lower (docs/tutorials/ascii_fhe_string.md:81-154) and the
// regex example's StringCiphertext = Vec<RadixCiphertext>, 4 blocks per char (examples/regex_engine/ciphertext.rs:4-22).
// The operations below are the compositions SURVEY.md Appendix B derives for BASELINE.json's configs, on strings of
// PUBLIC length (no padding): a char is an FheUint8 = 4 little-endian 2-bit blocks (integer/encryption.rs:69-83).
// Only decrypted results are comparable with the reference (there is nothing to compare ciphertext-wise).
#pragma once
#include "radix.h"

namespace tbh {

struct FheString {
    std::vector<Radix> chars;   // chars[i] = 4 blocks, little-endian
    size_t len() const { return chars.size(); }
};

class StringServerKey {
  public:
    explicit StringServerKey(Program &prog, bool packed_eq = false) : pg(prog), isk(prog), p(prog.params()), packed(packed_eq) {}

    // a char is an FheUint8: ceil(8 / log2(message_modulus)) little-endian blocks (integer/encryption.rs:69-83) -- 4 blocks of 2 bits for
    // the *_2_CARRY_* sets every config of BASELINE.json uses, 8 of 1 bit for MESSAGE_1, 3 (3 + 3 + 2 bits) for MESSAGE_3, 2 for MESSAGE_4
    static size_t blocks_per_char(const Params &q) {
        unsigned bits = 0;
        while ((uint64_t(1) << (bits + 1)) <= q.msg_mod) ++bits;
        if (bits == 0 || (uint64_t(1) << bits) != q.msg_mod) throw std::invalid_argument("strings: the message modulus must be a power of two >= 2");
        return (8 + bits - 1) / bits;
    }
    size_t blocks_per_char() const { return blocks_per_char(p); }
    uint64_t char_block(unsigned char ch, size_t b) const {
        unsigned bits = 0;
        while ((uint64_t(1) << (bits + 1)) <= p.msg_mod) ++bits;
        return (uint64_t(ch) >> (bits * b)) & (p.msg_mod - 1);
    }
    // the fused case conversion, the find index arithmetic and the null-padded model are written for 2-bit blocks (message modulus 4)
    void require_two_bit_blocks(const char *what) const {
        if (p.msg_mod != 4) throw std::invalid_argument(std::string(what) + ": implemented for message modulus 4 (2-bit blocks) only");
    }

    FheString input_string(size_t n_chars) {
        FheString s;
        for (size_t i = 0; i < n_chars; ++i) {
            Radix c;
            for (size_t b = 0; b < blocks_per_char(); ++b) c.push_back(pg.input());
            s.chars.push_back(c);
        }
        return s;
    }
    // trivial (clear) string, e.g. a clear pattern: server_key/mod.rs:684-721 per block
    FheString trivial_string(const std::string &clear) {
        FheString s;
        for (unsigned char ch : clear) {
            Radix c;
            for (size_t b = 0; b < blocks_per_char(); ++b) c.push_back(pg.create_trivial(char_block(ch, b)));
            s.chars.push_back(c);
        }
        return s;
    }

    static Radix concat(const FheString &s, size_t from, size_t n) {
        Radix r;
        for (size_t i = from; i < from + n; ++i) r.insert(r.end(), s.chars[i].begin(), s.chars[i].end());
        return r;
    }
    // big-endian integer view for lexicographic order: char 0 most significant (SURVEY.md App. B "Ordering")
    static Radix lex_radix(const FheString &s, size_t n) {
        Radix r;
        for (size_t i = n; i-- > 0;) r.insert(r.end(), s.chars[i].begin(), s.chars[i].end());
        return r;
    }

    // ---- eq / ne: config 1 ([32, 3, 1] PBS per level for 8 chars) -----------------------------------------------------
    // block equalities of two equal-length block ranges: reference form (one bivariate PBS per block, comparison.rs:10-33)
    // or the packed form (one PBS per pair of blocks, radix.h packed_block_equalities)
    std::vector<Ct> block_eqs(const Radix &x, const Radix &y) {
        return packed ? isk.packed_block_equalities(x, y) : isk.block_equalities(x, y);
    }
    BooleanBlock eq(const FheString &a, const FheString &b) {
        if (a.len() != b.len()) return pg.create_trivial(0);
        return isk.are_all_comparisons_block_true(block_eqs(concat(a, 0, a.len()), concat(b, 0, b.len())));
    }
    BooleanBlock ne(const FheString &a, const FheString &b) {
        if (a.len() != b.len()) return pg.create_trivial(1);
        return isk.unchecked_ne(concat(a, 0, a.len()), concat(b, 0, b.len()));
    }

    // ---- lexicographic comparison: config 5 ([256,128,...,1,1] for 128 chars) ----------------------------------------
    // sign of the common prefix; the public lengths break ties
    BooleanBlock cmp(const FheString &a, const FheString &b, bool want_less, bool or_equal) {
        const size_t n = std::min(a.len(), b.len());
        const bool a_shorter = a.len() < b.len(), same = a.len() == b.len();
        // result when the common prefix is equal
        const bool tie = same ? or_equal : (want_less ? a_shorter : !a_shorter);
        if (n == 0) return pg.create_trivial(tie ? 1 : 0);
        Ct sign = isk.unchecked_compare(lex_radix(a, n), lex_radix(b, n));
        return isk.map_sign_result(sign, [=](uint64_t x) {
            if (x == IntegerServerKey::IS_EQUAL) return tie;
            return want_less ? x == IntegerServerKey::IS_INFERIOR : x == IntegerServerKey::IS_SUPERIOR;
        });
    }
    BooleanBlock lt(const FheString &a, const FheString &b) { return cmp(a, b, true, false); }
    BooleanBlock le(const FheString &a, const FheString &b) { return cmp(a, b, true, true); }
    BooleanBlock gt(const FheString &a, const FheString &b) { return cmp(a, b, false, false); }
    BooleanBlock ge(const FheString &a, const FheString &b) { return cmp(a, b, false, true); }

    // multi-GPU split of a comparison: a rank's share is the sign block of its char range ...
    Ct compare_sign(const FheString &a, const FheString &b) {
        if (a.len() != b.len() || a.len() == 0) throw std::invalid_argument("compare_sign: equal, non-zero lengths required");
        return isk.unchecked_compare(lex_radix(a, a.len()), lex_radix(b, b.len()));
    }
    // ... and the gathered sign blocks (least significant range first) finish with the pairwise tree + the final map (comparator.rs:257-279,957-971)
    BooleanBlock finish_signs(std::vector<Ct> signs, bool want_less, bool or_equal) {
        Ct sign = isk.reduce_signs(std::move(signs));
        return isk.map_sign_result(sign, [=](uint64_t x) {
            if (x == IntegerServerKey::IS_EQUAL) return or_equal;
            return want_less ? x == IntegerServerKey::IS_INFERIOR : x == IntegerServerKey::IS_SUPERIOR;
        });
    }
    // multi-GPU split of find: every rank's (found, first index in its window range), ranges in ascending order -> the overall first match.
    //   level 1   first_r = [found_r and no earlier rank found]        (leveled prefix count of the flags, one PBS per rank)
    //             found   = [sum of the flags != 0]
    //   level 2   index digit b = sum_r (first_r ? index_r[b] : 0)     (one PBS per rank and digit; at most one term is non-zero)
    std::pair<BooleanBlock, Radix> combine_find(const std::vector<std::pair<BooleanBlock, Radix>> &parts) {
        if (parts.empty()) throw std::invalid_argument("combine_find: no parts");
        if (2 * parts.size() > p.total_mod()) throw std::invalid_argument("combine_find: too many parts for the message space");
        const size_t nb = parts[0].second.size();
        std::vector<Ct> flags, first;
        Ct before = pg.create_trivial(0);
        for (size_t r = 0; r < parts.size(); ++r) {
            flags.push_back(parts[r].first);
            // x = found_r + 2 * (#earlier ranks that found)  in [0, 15]
            Ct x = pg.unchecked_add(parts[r].first, pg.unchecked_scalar_mul(before, 2));
            first.push_back(r == 0 ? parts[r].first : pg.pbs(x, [](uint64_t v) { return uint64_t(v == 1); }));
            before = pg.unchecked_add(before, parts[r].first);
        }
        Ct found = isk.is_at_least_one_comparisons_block_true(flags);
        Radix index;
        for (size_t b = 0; b < nb; ++b) {
            Ct sum;
            for (size_t r = 0; r < parts.size(); ++r) {
                Ct y = pg.unchecked_add(pg.unchecked_scalar_mul(first[r], p.msg_mod), parts[r].second[b]);
                const uint64_t mm = p.msg_mod;
                Ct term = pg.pbs(y, [mm](uint64_t v) { return (v / mm) % 2 ? v % mm : uint64_t(0); });
                sum = r == 0 ? term : pg.unchecked_add(sum, term);
            }
            sum.degree = p.msg_mod - 1;       // at most one term is non-zero
            index.push_back(sum);
        }
        return {found, index};
    }

    // ---- contains / starts_with / ends_with / find: config 3 -------------------------------------------------------------
    // match flag of every window: block equalities for all (window, pattern position) pairs in one level, then one
    // AND tree per window ([15424, 1205, 241] for 256/16)
    // windows [w0, w1) only (w1 = npos: all) -- the multi-GPU split records just a rank's share of the windows
    std::vector<Ct> window_matches(const FheString &hay, const FheString &pat, size_t w0 = 0, size_t w1 = size_t(-1)) {
        std::vector<Ct> m;
        if (pat.len() > hay.len()) return m;
        const size_t W = hay.len() - pat.len() + 1;
        w1 = std::min(w1, W);
        std::vector<std::vector<Ct>> eqs;
        for (size_t w = w0; w < w1; ++w) eqs.push_back(block_eqs(concat(hay, w, pat.len()), concat(pat, 0, pat.len())));
        // the AND trees are scheduled by readiness, so the per-window reductions line up level by level
        for (auto &e : eqs) m.push_back(isk.are_all_comparisons_block_true(e));
        return m;
    }
    BooleanBlock contains(const FheString &hay, const FheString &pat) {
        if (pat.len() == 0) return pg.create_trivial(1);
        if (pat.len() > hay.len()) return pg.create_trivial(0);
        std::vector<Ct> m = window_matches(hay, pat);
        return isk.is_at_least_one_comparisons_block_true(m);
    }
    BooleanBlock starts_with(const FheString &s, const FheString &pat) {
        if (pat.len() > s.len()) return pg.create_trivial(0);
        if (pat.len() == 0) return pg.create_trivial(1);
        return isk.are_all_comparisons_block_true(block_eqs(concat(s, 0, pat.len()), concat(pat, 0, pat.len())));
    }
    BooleanBlock ends_with(const FheString &s, const FheString &pat) {
        if (pat.len() > s.len()) return pg.create_trivial(0);
        if (pat.len() == 0) return pg.create_trivial(1);
        return isk.are_all_comparisons_block_true(block_eqs(concat(s, s.len() - pat.len(), pat.len()), concat(pat, 0, pat.len())));
    }
    // find: (found, index of the first match as a radix of ceil(log4(W)) blocks; 0 when not found, like the
    // regex example's "no match" convention of returning a boolean separately)
    std::pair<BooleanBlock, Radix> find(const FheString &hay, const FheString &pat) { return find_range(hay, pat, 0, size_t(-1)); }
    // rfind: the LAST match -- the same circuit over the windows in descending order
    std::pair<BooleanBlock, Radix> rfind(const FheString &hay, const FheString &pat) { return find_range(hay, pat, 0, size_t(-1), true); }
    // number of radix blocks of a window index
    static size_t index_blocks(size_t hay_len, size_t pat_len) {
        const size_t W = pat_len <= hay_len ? hay_len - pat_len + 1 : 0;
        size_t idx_blocks = 1;
        while ((size_t(1) << (2 * idx_blocks)) < std::max<size_t>(W, 1)) ++idx_blocks;
        return idx_blocks;
    }
    // the same over the windows [w0, w1) only (the multi-GPU split): first match inside the range, reported as its GLOBAL window index
    std::pair<BooleanBlock, Radix> find_range(const FheString &hay, const FheString &pat, size_t w0, size_t w1, bool last = false) {
        require_two_bit_blocks("find");
        const size_t W_all = pat.len() <= hay.len() ? hay.len() - pat.len() + 1 : 0;
        const size_t idx_blocks = index_blocks(hay.len(), pat.len());
        Radix zero_idx;
        for (size_t b = 0; b < idx_blocks; ++b) zero_idx.push_back(pg.create_trivial(0));
        if (W_all == 0) return {pg.create_trivial(0), zero_idx};
        w1 = std::min(w1, W_all);
        if (w0 >= w1) return {pg.create_trivial(0), zero_idx};
        if (pat.len() == 0) {                // the empty pattern matches everywhere: first window 0, last window W_all - 1 (public)
            if (!last) return {pg.create_trivial(w0 == 0 ? 1 : 0), zero_idx};
            Radix idx;
            for (size_t b = 0; b < idx_blocks; ++b) idx.push_back(pg.create_trivial(w1 == W_all ? ((W_all - 1) >> (2 * b)) & 3 : 0));
            return {pg.create_trivial(w1 == W_all ? 1 : 0), idx};
        }
        // local window w is global window w0 + w (w1 - 1 - w for rfind)
        std::vector<Ct> m = window_matches(hay, pat, w0, w1);
        if (last) std::reverse(m.begin(), m.end());
        return first_true(m, [=](size_t w) { return last ? w1 - 1 - w : w0 + w; }, idx_blocks);
    }
    // (found, index) of the FIRST true flag of m, reported as global(w) in idx_blocks radix digits (zero when nothing is true)
    template <class Global>
    std::pair<BooleanBlock, Radix> first_true(const std::vector<Ct> &m, Global global, size_t idx_blocks) {
        require_two_bit_blocks("find");
        const size_t W = m.size();
        // one-hot first match in THREE levels instead of a log-depth prefix OR.  Windows are cut into blocks of 14:
        //   level 1   any_k   = [sum of the block's match flags != 0]                (leveled sum of <= 14 booleans + 1 PBS per block)
        //   level 2   before_k = [sum_{k' < k} any_k' != 0]                         (leveled prefix sums, chunks of <= 15, 1 PBS per block)
        //   level 3   first_w = [ (#matches before w inside its block) + before_k + (1 - m_w) == 0 ]
        //             -- the in-block prefix count is a leveled sum (<= 13), so the argument stays <= 15; the same PBS multiplies by
        //             the window's index digit (one PBS per window and non-zero digit).
        // found = OR of the blocks' any flags (same count-tree as contains).
        constexpr size_t BLK = 14;
        const size_t n_blk = (W + BLK - 1) / BLK;
        std::vector<Ct> any(n_blk);
        for (size_t k = 0; k < n_blk; ++k) {
            Ct sum = m[k * BLK];
            for (size_t i = k * BLK + 1; i < std::min(W, (k + 1) * BLK); ++i) sum = pg.unchecked_add(sum, m[i]);
            any[k] = pg.pbs(sum, [](uint64_t x) { return uint64_t(x != 0); });
        }
        // before_k: "some earlier block matched".  A prefix sum of booleans must stay <= total_mod - 1 = 15 (the next value would reach
        // the padding bit and the LUT would return -f(0)), so a cleaned flag is carried forward from group to group: the first group
        // sums 15 blocks, every later group its carry + 14 blocks (one extra PBS per group, still the same level for the first group).
        std::vector<Ct> before(n_blk);
        before[0] = pg.create_trivial(0);
        {
            const size_t max_sum = p.total_mod() - 1;
            Ct carry = pg.create_trivial(0);   // cleaned "matched before this group of blocks"
            for (size_t g0 = 0; g0 < n_blk;) {
                const size_t g1 = std::min(n_blk, g0 + (g0 == 0 ? max_sum : max_sum - 1));
                Ct run = carry;
                for (size_t k = g0; k < g1; ++k) {
                    if (k > 0) before[k] = pg.pbs(run, [](uint64_t x) { return uint64_t(x != 0); });
                    run = pg.unchecked_add(run, any[k]);
                }
                if (g1 < n_blk) carry = pg.pbs(run, [](uint64_t x) { return uint64_t(x != 0); });
                g0 = g1;
            }
        }
        // y_w = (matches before w in the block) + before_k + 1 - m_w   in [0, 15];  first_w = [y_w == 0] is never materialised: the
        // index digits are selected straight from y_w (one LUT per non-zero digit value), which saves a tree level and W PBS
        std::vector<Ct> y(W);
        for (size_t w = 0; w < W; ++w) {
            const size_t k = w / BLK;
            Ct t = pg.unchecked_scalar_add(pg.unchecked_scalar_mul(m[w], uint64_t(-1)), 1);
            t.degree = 1;
            for (size_t i = k * BLK; i < w; ++i) t = pg.unchecked_add(t, m[i]);
            y[w] = pg.unchecked_add(t, before[k]);
        }
        Ct found = isk.is_at_least_one_comparisons_block_true(any);
        // index digit b = sum_w ((w >> 2b) & 3) * first_w: select with a LUT (clean, noise NOMINAL), then sum in
        // chunks of 15 with a cleaning PBS, exactly the chunking of scalar_comparison.rs:155-170
        const size_t max_value = p.total_mod() - 1;
        Radix index;
        for (size_t b = 0; b < idx_blocks; ++b) {
            std::vector<Ct> terms;
            for (size_t w = 0; w < W; ++w) {
                const uint64_t digit = (global(w) >> (2 * b)) & 3;
                if (digit == 0) continue;
                terms.push_back(pg.pbs(y[w], [digit](uint64_t x) { return x == 0 ? digit : uint64_t(0); }));
            }
            if (terms.empty()) { index.push_back(pg.create_trivial(0)); continue; }
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
            index.push_back(terms[0]);
        }
        return {found, index};
    }

    // ---- case conversion: config 4 --------------------------------------------------------------------------------------------
    // The tutorial circuit (ascii_fhe_string.md:88-94: gt(64) & lt(91), cast, *32, +) costs 19 PBS per char; only the
    // decrypted result has to match, so the fused form is used: with hi = b3*4+b2 and lo = b1*4+b0 (leveled packing),
    //   class(hi) in {0, 1: hi == H, 2: hi == H+1},  range(lo) = [lo >= 1] + 2*[lo <= 10],
    //   is_letter = LUT(class*4 + range),  b2' = LUT(b2*4 + is_letter) = b2 +/- 2   (bit 5 toggles; no carry can occur)
    // = 4 PBS per char in 3 levels.  H = 4 for upper-case input ('A'..'Z' = 0x41..0x5A), 6 for lower-case.
    Ct is_letter_of_case(const Radix &c, bool upper) {
        const uint64_t H = upper ? 4 : 6;
        Ct hi = pg.unchecked_add(pg.unchecked_scalar_mul(c[3], p.msg_mod), c[2]);
        Ct lo = pg.unchecked_add(pg.unchecked_scalar_mul(c[1], p.msg_mod), c[0]);
        Ct cls = pg.pbs(hi, [H](uint64_t x) { return x == H ? uint64_t(1) : (x == H + 1 ? uint64_t(2) : uint64_t(0)); });
        Ct rng = pg.pbs(lo, [](uint64_t x) { return uint64_t(x >= 1) + 2 * uint64_t(x <= 10); });
        return pg.pbs_bivariate(cls, rng, [](uint64_t k, uint64_t r) {
            return uint64_t((k == 1 && (r & 1)) || (k == 2 && (r & 2)));
        });
    }
    FheString change_case(const FheString &s, bool to_lower) {
        require_two_bit_blocks("case conversion");
        FheString out = s;
        for (size_t i = 0; i < s.len(); ++i) {
            Ct flag = is_letter_of_case(s.chars[i], /*upper=*/to_lower);
            out.chars[i][2] = pg.pbs_bivariate(s.chars[i][2], flag, [to_lower](uint64_t b2, uint64_t f) {
                if (!(f & 1)) return b2;
                return to_lower ? (b2 | 2) : (b2 & 1);   // set / clear bit 5 of the char (bit 1 of block 2)
            });
        }
        return out;
    }
    FheString to_lowercase(const FheString &s) { return change_case(s, true); }
    FheString to_uppercase(const FheString &s) { return change_case(s, false); }
    BooleanBlock eq_ignore_case(const FheString &a, const FheString &b) {
        if (a.len() != b.len()) return pg.create_trivial(0);
        return eq(to_lowercase(a), to_lowercase(b));
    }

    Program &pg;
    IntegerServerKey isk;
    Params p;
    bool packed;   // use packed block equalities in eq / contains / find / starts_with / ends_with / eq_ignore_case
};

}  // namespace tbh
