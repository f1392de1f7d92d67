// host/program.h -- host-side mirror of the shortint layer that sits on the KS-PBS path, restructured
// as a level-synchronous batch scheduler.
//
// Reference interfaces mirrored (tfhe/src/shortint/):
//   Ciphertext{ct, degree, noise_level, ..}            ciphertext/mod.rs:263-270   -> Ct (lazy linear expression + metadata)
//   ServerKey::generate_lookup_table                   server_key/mod.rs:383-399   -> LutRegistry::get
//   fill_accumulator                                   engine/mod.rs:72-128        -> LutRegistry::fill_accumulator
//   generate_lookup_table_bivariate                    server_key/bivariate_pbs.rs:71-97
//   apply_lookup_table / keyswitch_programmable_bootstrap_assign   server_key/mod.rs:457-476,783-857 -> Program::pbs
//   trivial_pbs_assign                                 server_key/mod.rs:763-781   -> Program::pbs on a trivial Ct (host side)
//   unchecked_add / scalar_mul / scalar_add / lwe sub  server_key/add.rs:520-524, scalar_mul.rs:520-536,
//                                                      scalar_add.rs:211-218, lwe_linear_algebra.rs:703
//   create_trivial                                     server_key/mod.rs:684-721
//
// The reference runs one PBS per rayon task (integer/server_key/radix_parallel/*.rs).  Here an operation
// is first *recorded*: leveled ops compose lazily into linear expressions over "real" arena slots, every
// apply_lookup_table becomes a job in the level after its operand is ready, and the whole tree is then
// executed as  [one leveled-op launch + one keyswitch launch + one PBS launch] per level.
// Nothing in this file touches CUDA; it is plain C++ and is unit-tested without a GPU.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace tbh {

struct Params {
    uint32_t lwe_dim, glwe_dim, poly_size, pbs_base_log, pbs_level, ks_base_log, ks_level, grouping_factor, msg_mod, carry_mod;
    uint32_t total_mod() const { return msg_mod * carry_mod; }
    uint64_t delta() const { return (uint64_t(1) << 63) / total_mod(); }
    size_t big_len() const { return size_t(glwe_dim) * poly_size + 1; }
    size_t lut_len() const { return size_t(glwe_dim + 1) * poly_size; }
};

// ---- lookup tables ---------------------------------------------------------------------------------
struct Lut {
    std::vector<uint64_t> table;  // f(0..total_mod-1)
    uint64_t degree;              // max f (engine/mod.rs:113)
};

class LutRegistry {
  public:
    explicit LutRegistry(const Params &p) : p_(p) {}

    uint32_t get(const std::function<uint64_t(uint64_t)> &f) {
        std::vector<uint64_t> t(p_.total_mod());
        for (uint32_t i = 0; i < p_.total_mod(); ++i) t[i] = f(i);
        auto it = index_.find(t);
        if (it != index_.end()) return it->second;
        Lut l;
        l.table = t;
        l.degree = *std::max_element(t.begin(), t.end());
        luts_.push_back(l);
        const uint32_t id = uint32_t(luts_.size() - 1);
        index_[t] = id;
        return id;
    }
    // bivariate_pbs.rs:84-89: f'(x) = f((x / msg_mod) % msg_mod, (x % msg_mod) % msg_mod)
    uint32_t get_bivariate(const std::function<uint64_t(uint64_t, uint64_t)> &f) {
        const uint64_t m = p_.msg_mod;
        return get([&](uint64_t x) { return f((x / m) % m, (x % m) % m); });
    }
    const Lut &lut(uint32_t id) const { return luts_[id]; }
    size_t size() const { return luts_.size(); }

    // engine/mod.rs:94-127: GLWE accumulator, mask = 0, body boxes of N/total_mod, first half-box negated,
    // rotate-left by half a box.
    void fill_accumulator(uint32_t id, uint64_t *acc) const {
        const size_t N = p_.poly_size, k = p_.glwe_dim, sup = p_.total_mod(), box = N / sup, half = box / 2;
        const uint64_t delta = p_.delta();
        std::fill(acc, acc + k * N, uint64_t(0));
        std::vector<uint64_t> tmp(N);
        for (size_t i = 0; i < sup; ++i)
            for (size_t j = 0; j < box; ++j) tmp[i * box + j] = luts_[id].table[i] * delta;
        for (size_t j = 0; j < half; ++j) tmp[j] = uint64_t(0) - tmp[j];
        for (size_t j = 0; j < N; ++j) acc[k * N + j] = tmp[(j + half) % N];
    }
    // server_key/mod.rs:763-781 on a plaintext body
    uint64_t trivial_pbs(uint32_t id, uint64_t body) const {
        const uint64_t sup = p_.total_mod(), delta = p_.delta();
        uint64_t v = body / delta;
        if (v >= sup) return uint64_t(0) - luts_[id].table[v % sup] * delta;  // padding bit set
        return luts_[id].table[v] * delta;
    }

  private:
    Params p_;
    std::vector<Lut> luts_;
    std::map<std::vector<uint64_t>, uint32_t> index_;
};

// ---- ciphertext handle -------------------------------------------------------------------------------
// A shortint::Ciphertext whose LWE data is the linear expression  sum_k coef_k * arena[slot_k]  (+ body on
// the body word).  No terms => trivial ciphertext (mask 0), exactly `is_trivial()` of ciphertext/mod.rs:371-374.
struct Ct {
    std::vector<std::pair<uint32_t, int64_t>> terms;
    uint64_t body = 0;
    uint64_t degree = 0;
    uint64_t noise = 0;    // NoiseLevel (ciphertext/mod.rs), NOMINAL = 1
    int ready = -1;        // index of the last level whose PBS outputs this expression reads (-1: inputs only)
    bool is_trivial() const { return terms.empty(); }
};

struct LinInstrH { uint32_t out_slot, term_begin, term_end; uint64_t body_add; };
struct LinTermH { uint32_t slot; int64_t coef; };
struct PbsJobH { uint32_t in_slot, out_slot, lut; };
struct Level {
    std::vector<LinInstrH> lin;   // executed before the level's PBS
    std::vector<PbsJobH> pbs;
};

class Program {
  public:
    Program(const Params &p, LutRegistry &luts) : p_(p), luts_(luts) {}

    const Params &params() const { return p_; }
    LutRegistry &luts() { return luts_; }

    // a fresh input ciphertext (client_side.rs:120-127: degree = msg_mod - 1, noise NOMINAL)
    Ct input() { return input(p_.msg_mod - 1, 1); }
    Ct input(uint64_t degree, uint64_t noise) {
        if (!levels_.empty() || n_slots_ != n_inputs_) throw std::logic_error("inputs must be declared first");
        Ct c;
        c.terms.push_back({n_slots_++, 1});
        ++n_inputs_;
        c.degree = degree; c.noise = noise;
        return c;
    }
    // server_key/mod.rs:684-721
    Ct create_trivial(uint64_t value) const { return unchecked_create_trivial(value % p_.msg_mod); }
    Ct unchecked_create_trivial(uint64_t value) const {
        Ct c; c.body = value * p_.delta(); c.degree = value; c.noise = 0; return c;
    }

    // ---- leveled ops (lazy) ---------------------------------------------------------------------------
    static void add_terms(Ct &dst, const Ct &src, int64_t k) {
        for (auto &t : src.terms) {
            bool found = false;
            for (auto &d : dst.terms) if (d.first == t.first) { d.second += t.second * k; found = true; break; }
            if (!found) dst.terms.push_back({t.first, t.second * k});
        }
        dst.terms.erase(std::remove_if(dst.terms.begin(), dst.terms.end(), [](auto &t) { return t.second == 0; }), dst.terms.end());
        dst.body += src.body * uint64_t(k);
        dst.ready = std::max(dst.ready, src.ready);
    }
    // add.rs:520-524
    Ct unchecked_add(const Ct &a, const Ct &b) const {
        Ct r = a; add_terms(r, b, 1); r.degree = a.degree + b.degree; r.noise = a.noise + b.noise; return r;
    }
    // lwe_linear_algebra.rs:703 (true LWE subtraction as used by comparator.rs:213); degree is the caller's business
    Ct lwe_sub(const Ct &a, const Ct &b) const {
        Ct r = a; add_terms(r, b, -1); r.noise = a.noise + b.noise; return r;
    }
    // scalar_mul.rs:520-536
    Ct unchecked_scalar_mul(const Ct &a, uint64_t k) const {
        Ct r; r.ready = a.ready; add_terms(r, a, int64_t(k)); r.degree = a.degree * k; r.noise = a.noise * k; return r;
    }
    // scalar_add.rs:211-218
    Ct unchecked_scalar_add(const Ct &a, uint64_t v) const {
        Ct r = a; r.body += v * p_.delta(); r.degree = a.degree + v; return r;
    }
    // lwe_linear_algebra.rs:384 (plaintext sub, comparator.rs:226-238)
    Ct plaintext_sub(const Ct &a, uint64_t v) const { Ct r = a; r.body -= v * p_.delta(); return r; }
    // bivariate_pbs.rs:176-178: lhs * msg_mod + rhs
    Ct pack_bivariate(const Ct &lhs, const Ct &rhs) const {
        return unchecked_add(unchecked_scalar_mul(lhs, p_.msg_mod), rhs);
    }

    // ---- apply_lookup_table (server_key/mod.rs:457-476 -> 783-857) ---------------------------------------------
    Ct pbs(const Ct &a, uint32_t lut) {
        const Lut &l = luts_.lut(lut);
        // the LUT only covers [0, total_mod): a larger value reaches the padding bit and comes back as -f(x - total_mod)
        // (max_degree of shortint/server_key/mod.rs:91-103: msg_mod * carry_mod - 1)
        if (a.degree >= p_.total_mod())
            throw std::logic_error("apply_lookup_table on a ciphertext of degree " + std::to_string(a.degree) + " >= message space " +
                                   std::to_string(p_.total_mod()));
        if (a.is_trivial()) {  // mod.rs:788-791 trivial short cut: pure table lookup, no crypto
            Ct r; r.body = luts_.trivial_pbs(lut, a.body); r.degree = l.degree; r.noise = 0; r.ready = a.ready;
            ++n_trivial_pbs_;
            return r;
        }
        const int lv = a.ready + 1;
        if (int(levels_.size()) <= lv) levels_.resize(lv + 1);
        const uint32_t in = materialize(a, lv);
        const uint32_t out = n_slots_++;
        levels_[lv].pbs.push_back({in, out, lut});
        Ct r; r.terms.push_back({out, 1}); r.degree = l.degree; r.noise = 1; r.ready = lv;  // mod.rs:855-856
        return r;
    }
    Ct pbs(const Ct &a, const std::function<uint64_t(uint64_t)> &f) { return pbs(a, luts_.get(f)); }
    // bivariate_pbs.rs:167-182
    Ct pbs_bivariate(const Ct &lhs, const Ct &rhs, uint32_t lut) { return pbs(pack_bivariate(lhs, rhs), lut); }
    Ct pbs_bivariate(const Ct &lhs, const Ct &rhs, const std::function<uint64_t(uint64_t, uint64_t)> &f) {
        return pbs(pack_bivariate(lhs, rhs), luts_.get_bivariate(f));
    }

    // mark a result: it is materialised into an arena slot that the executor downloads
    void output(const Ct &a) {
        const int lv = a.ready + 1;
        if (!(a.terms.size() == 1 && a.terms[0].second == 1 && a.body == 0) && int(levels_.size()) <= lv) levels_.resize(lv + 1);
        outputs_.push_back(materialize(a, lv));
        output_meta_.push_back(a);
    }

    // ---- accessors for the executor / tests ----------------------------------------------------------------------
    uint32_t n_slots() const { return n_slots_; }
    uint32_t n_inputs() const { return n_inputs_; }
    const std::vector<Level> &levels() const { return levels_; }
    const std::vector<LinTermH> &terms() const { return terms_; }
    const std::vector<uint32_t> &outputs() const { return outputs_; }
    const std::vector<Ct> &output_meta() const { return output_meta_; }
    size_t n_pbs() const { size_t s = 0; for (auto &l : levels_) s += l.pbs.size(); return s; }
    size_t n_trivial_pbs() const { return n_trivial_pbs_; }
    std::vector<size_t> level_widths() const {
        std::vector<size_t> w;
        for (auto &l : levels_) if (!l.pbs.empty()) w.push_back(l.pbs.size());
        return w;
    }

  private:
    // returns the arena slot holding `a`, emitting a leveled instruction at level `lv` if `a` is not a bare slot
    uint32_t materialize(const Ct &a, int lv) {
        if (a.terms.size() == 1 && a.terms[0].second == 1 && a.body == 0) return a.terms[0].first;
        if (int(levels_.size()) <= lv) levels_.resize(lv + 1);
        LinInstrH ins;
        ins.out_slot = n_slots_++;
        ins.term_begin = uint32_t(terms_.size());
        for (auto &t : a.terms) terms_.push_back({t.first, t.second});
        ins.term_end = uint32_t(terms_.size());
        ins.body_add = a.body;
        levels_[lv].lin.push_back(ins);
        return ins.out_slot;
    }

    Params p_;
    LutRegistry &luts_;
    uint32_t n_slots_ = 0, n_inputs_ = 0;
    std::vector<Level> levels_;
    std::vector<LinTermH> terms_;
    std::vector<uint32_t> outputs_;
    std::vector<Ct> output_meta_;
    size_t n_trivial_pbs_ = 0;
};

}  // namespace tbh
