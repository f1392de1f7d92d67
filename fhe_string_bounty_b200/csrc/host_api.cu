// host_api.cu -- C ABI of the host layer: record an integer / string operation as a level-batched program
// (host/program.h, host/radix.h, host/strings.h), inspect it without a GPU, and execute it on a context:
// per level ONE leveled-op launch + ONE keyswitch launch + ONE PBS launch over every independent block.
// This is the restructuring of the reference's per-block rayon fan-out
// (integer/server_key/radix_parallel/*.rs -> shortint apply_lookup_table) that north_star asks for.
#include "ctx.h"
#include "host/padded.h"

#include <memory>

using tbc::DevBuf;
using tbc::DeviceGuard;
using tbc::fail;

struct tfhe_b200_program {
    tbh::Params p;
    std::unique_ptr<tbh::LutRegistry> luts;
    std::unique_ptr<tbh::Program> prog;
    std::string op;
    // device copies, valid for the context with id `owner_id` on CUDA device `device` (ids are unique per process, so a destroyed
    // context whose address is reused is never mistaken for the owner; the buffers are plain device memory and outlive the context)
    uint64_t owner_id = 0;
    int device = -1;
    DevBuf d_luts, d_lin, d_terms, d_pbs_in, d_pbs_out, d_pbs_lut, d_arena, d_out_slots, d_out_rows;
    std::vector<uint32_t> lin_off, pbs_off;   // per level offsets
    float last_ms = 0.f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

tbh::Params to_host_params(const tfhe_b200_params &q) {
    return tbh::Params{q.lwe_dim, q.glwe_dim, q.poly_size, q.pbs_base_log, q.pbs_level, q.ks_base_log, q.ks_level,
                       q.grouping_factor, q.msg_mod, q.carry_mod};
}

tbh::Radix input_radix(tbh::Program &pg, size_t n) {
    tbh::Radix r;
    for (size_t i = 0; i < n; ++i) r.push_back(pg.input());
    return r;
}

void output_radix(tbh::Program &pg, const tbh::Radix &r) { for (auto &b : r) pg.output(b); }
void output_string(tbh::Program &pg, const tbh::FheString &s) { for (auto &c : s.chars) output_radix(pg, c); }

// records `op`; returns false with a message when the op / arguments are unknown
bool record(tfhe_b200_program &h, const std::string &op_in, const uint64_t *a, size_t na, const char *clear, std::string &err) {
    tbh::Program &pg = *h.prog;
    tbh::IntegerServerKey isk(pg);
    // "<op>_packed": same operation with packed block equalities (one PBS per pair of blocks, host/radix.h)
    std::string op = op_in;
    bool packed = false;
    if (op.size() > 7 && op.compare(op.size() - 7, 7, "_packed") == 0) { packed = true; op = op.substr(0, op.size() - 7); }
    tbh::StringServerKey ssk(pg, packed);
    if (packed && op == "radix_eq") {
        if (na < 1) { err = "radix_eq_packed: expected 1 integer argument"; return false; }
        tbh::Radix x = input_radix(pg, a[0]), y = input_radix(pg, a[0]);
        pg.output(isk.unchecked_eq_packed(x, y));
        return true;
    }
    auto need = [&](size_t n) { if (na < n) { err = op + ": expected " + std::to_string(n) + " integer arguments"; return false; } return true; };
    const std::string cl = clear ? clear : "";

    // ---- shortint level: one apply_lookup_table per input, LUT given as a table in `a` --------------------------------------
    if (op == "shortint_apply_lut") {            // a = {n_cts, f(0), ..., f(total_mod-1)}
        if (!need(1 + h.p.total_mod())) return false;
        std::vector<uint64_t> t(a + 1, a + 1 + h.p.total_mod());
        std::vector<tbh::Ct> in;
        for (size_t i = 0; i < a[0]; ++i) in.push_back(pg.input(h.p.total_mod() - 1, 1));
        for (auto &c : in) pg.output(pg.pbs(c, [t](uint64_t x) { return t[x]; }));
        return true;
    }
    if (op == "shortint_bivariate_lut") {        // a = {n_pairs, f(0,0), f(0,1), ..., f(m-1,m-1)}  row-major in (lhs, rhs)
        const size_t m = h.p.msg_mod;
        if (!need(1 + m * m)) return false;
        std::vector<uint64_t> t(a + 1, a + 1 + m * m);
        std::vector<tbh::Ct> l, r;
        for (size_t i = 0; i < a[0]; ++i) l.push_back(pg.input());
        for (size_t i = 0; i < a[0]; ++i) r.push_back(pg.input());
        for (size_t i = 0; i < a[0]; ++i) pg.output(pg.pbs_bivariate(l[i], r[i], [t, m](uint64_t x, uint64_t y) { return t[x * m + y]; }));
        return true;
    }
    // ---- integer radix level: a = {n_blocks[, scalar]} ---------------------------------------------------------------------------------
    if (op.rfind("radix_", 0) == 0) {
        if (!need(1)) return false;
        const size_t nb = a[0];
        const std::string f = op.substr(6);
        if (f == "scalar_eq" || f == "scalar_lt" || f == "scalar_gt") {
            if (!need(2)) return false;
            tbh::Radix x = input_radix(pg, nb);
            pg.output(f == "scalar_eq" ? isk.unchecked_scalar_eq(x, a[1]) : f == "scalar_lt" ? isk.unchecked_scalar_lt(x, a[1]) : isk.unchecked_scalar_gt(x, a[1]));
            return true;
        }
        // default (propagating) forms on operands whose blocks carry up to a[1] (e.g. 6 after one unchecked_add): what the reference's
        // comparison tests exercise (radix_parallel/tests_cases_comparisons.rs:81-97)
        if (f.rfind("default_", 0) == 0 || f == "full_propagate") {
            if (!need(2)) return false;
            if (a[1] >= h.p.total_mod()) { err = op + ": block degree must stay below the message space"; return false; }
            auto dirty = [&](size_t n) { tbh::Radix r; for (size_t i = 0; i < n; ++i) r.push_back(pg.input(a[1], 2)); return r; };
            tbh::Radix x = dirty(nb);
            if (f == "full_propagate") { output_radix(pg, isk.full_propagate_any_degree(x)); return true; }
            tbh::Radix y = dirty(nb);
            x = isk.cleaned(x); y = isk.cleaned(y);
            const std::string g = f.substr(8);
            if (g == "eq") pg.output(isk.unchecked_eq(x, y));
            else if (g == "ne") pg.output(isk.unchecked_ne(x, y));
            else if (g == "lt") pg.output(isk.unchecked_lt(x, y));
            else if (g == "le") pg.output(isk.unchecked_le(x, y));
            else if (g == "gt") pg.output(isk.unchecked_gt(x, y));
            else if (g == "ge") pg.output(isk.unchecked_ge(x, y));
            else { err = "unknown radix op: " + op; return false; }
            return true;
        }
        if (f == "if_then_else") {
            tbh::Ct cond = pg.input(1, 1);
            tbh::Radix x = input_radix(pg, nb), y = input_radix(pg, nb);
            output_radix(pg, isk.if_then_else(cond, x, y));
            return true;
        }
        tbh::Radix x = input_radix(pg, nb), y = input_radix(pg, nb);
        if (f == "eq") pg.output(isk.unchecked_eq(x, y));
        else if (f == "ne") pg.output(isk.unchecked_ne(x, y));
        else if (f == "lt") pg.output(isk.unchecked_lt(x, y));
        else if (f == "le") pg.output(isk.unchecked_le(x, y));
        else if (f == "gt") pg.output(isk.unchecked_gt(x, y));
        else if (f == "ge") pg.output(isk.unchecked_ge(x, y));
        else if (f == "add") output_radix(pg, isk.add(x, y));
        else { err = "unknown radix op: " + op; return false; }
        return true;
    }
    // ---- boolean reductions used by the multi-GPU split: a = {n_blocks}; inputs are boolean blocks -------------------------------------------
    if (op == "bool_all_true" || op == "bool_any_true") {
        if (!need(1)) return false;
        std::vector<tbh::Ct> b;
        for (size_t i = 0; i < a[0]; ++i) b.push_back(pg.input(1, 1));
        pg.output(op == "bool_all_true" ? isk.are_all_comparisons_block_true(b) : isk.is_at_least_one_comparisons_block_true(b));
        return true;
    }
    // final LUT after an all-reduce(SUM) of per-GPU boolean blocks: a = {n_summed, want_all}
    if (op == "bool_sum_finish") {
        if (!need(2)) return false;
        const uint64_t n = a[0];
        tbh::Ct s = pg.input(n, n);
        pg.output(a[1] ? pg.pbs(s, [n](uint64_t x) { return uint64_t(x == n); }) : pg.pbs(s, [](uint64_t x) { return uint64_t(x != 0); }));
        return true;
    }
    // ---- multi-GPU finishing programs: inputs are what the ranks exchanged -------------------------------------------------
    if (op == "signs_finish") {                  // a = {count, want_less, or_equal}; inputs: count sign blocks, least significant range first
        if (!need(3)) return false;
        if (a[0] < 1) { err = "signs_finish: needs at least one sign block"; return false; }
        std::vector<tbh::Ct> signs;
        for (uint64_t r = 0; r < a[0]; ++r) signs.push_back(pg.input(2, 1));
        pg.output(ssk.finish_signs(signs, a[1] != 0, a[2] != 0));
        return true;
    }
    if (op == "find_combine") {                  // a = {count, index blocks}; inputs: count x (found flag, index radix), ascending window ranges
        if (!need(2)) return false;
        if (a[0] < 1 || 2 * a[0] > h.p.total_mod() || a[1] < 1) { err = "find_combine: 1..total_mod/2 parts, at least one index block"; return false; }
        std::vector<std::pair<tbh::Ct, tbh::Radix>> parts;
        for (uint64_t r = 0; r < a[0]; ++r) {
            tbh::Ct f = pg.input(1, 1);
            tbh::Radix idx;
            for (uint64_t b = 0; b < a[1]; ++b) idx.push_back(pg.input(h.p.msg_mod - 1, 1));
            parts.push_back({f, idx});
        }
        auto r = ssk.combine_find(parts);
        pg.output(r.first);
        output_radix(pg, r.second);
        return true;
    }
    // ---- many independent string pairs in one program (throughput mode): a = {len_a, len_b, count}; inputs are
    //      count x (string a, string b); one boolean per pair.  op = string_eq_many / string_lt_many / string_contains_many
    if (op.size() > 5 && op.rfind("string_", 0) == 0 && op.compare(op.size() - 5, 5, "_many") == 0) {
        if (!need(3)) return false;
        const std::string f = op.substr(7, op.size() - 12);
        std::vector<std::pair<tbh::FheString, tbh::FheString>> in;
        for (uint64_t k = 0; k < a[2]; ++k) {
            tbh::FheString s = ssk.input_string(a[0]);
            tbh::FheString t = ssk.input_string(a[1]);
            in.push_back({s, t});
        }
        for (auto &st : in) {
            if (f == "eq") pg.output(ssk.eq(st.first, st.second));
            else if (f == "ne") pg.output(ssk.ne(st.first, st.second));
            else if (f == "lt") pg.output(ssk.lt(st.first, st.second));
            else if (f == "le") pg.output(ssk.le(st.first, st.second));
            else if (f == "contains") pg.output(ssk.contains(st.first, st.second));
            else { err = "unknown batched string op: " + op; return false; }
        }
        return true;
    }
    // ---- null-padded strings (secret length, host/padded.h): a = {capacity_a[, capacity_b | count]}; a clear pattern comes through `clear` -----
    if (op.rfind("pstring_", 0) == 0) {
        if (!need(1)) return false;
        tbh::PaddedStringServerKey psk(pg);
        const std::string f = op.substr(8);
        tbh::FheString s = ssk.input_string(a[0]);
        if (f == "len") { output_radix(pg, psk.len(s)); return true; }
        if (f == "is_empty") { pg.output(psk.is_empty(s)); return true; }
        if (f == "trim_start") { output_string(pg, psk.trim_start(s)); return true; }
        if (f == "trim_end") { output_string(pg, psk.trim_end(s)); return true; }
        if (f == "trim") { output_string(pg, psk.trim(s)); return true; }
        if (f == "strip_prefix" || f == "strip_suffix") {
            if (!clear) { err = op + ": needs a clear pattern"; return false; }
            auto r = f == "strip_prefix" ? psk.strip_prefix(s, cl) : psk.strip_suffix(s, cl);
            pg.output(r.first);
            output_string(pg, r.second);
            return true;
        }
        if (f == "repeat") { if (!need(2)) return false; output_string(pg, psk.repeat(s, a[1])); return true; }
        if (!need(2)) return false;
        tbh::FheString t = ssk.input_string(a[1]);
        if (f == "eq") pg.output(psk.eq(s, t));
        else if (f == "ne") pg.output(psk.ne(s, t));
        else if (f == "lt") pg.output(psk.lt(s, t));
        else if (f == "le") pg.output(psk.le(s, t));
        else if (f == "gt") pg.output(psk.gt(s, t));
        else if (f == "ge") pg.output(psk.ge(s, t));
        else if (f == "contains") pg.output(psk.contains(s, t));
        else if (f == "starts_with") pg.output(psk.starts_with(s, t));
        else if (f == "ends_with") pg.output(psk.ends_with(s, t));
        else if (f == "concat") output_string(pg, psk.concat(s, t));
        else if (f == "find" || f == "rfind") { auto r = psk.find(s, t, f == "rfind"); pg.output(r.first); output_radix(pg, r.second); }
        else { err = "unknown padded string op: " + op; return false; }
        return true;
    }
    // ---- strings: a = {len_a[, len_b]}; with a non-empty `clear` the second operand is a clear (trivial) string ----------------------------------
    if (op.rfind("string_", 0) == 0) {
        if (!need(1)) return false;
        const std::string f = op.substr(7);
        tbh::FheString s = ssk.input_string(a[0]);
        if (f == "to_lowercase") { output_string(pg, ssk.to_lowercase(s)); return true; }
        if (f == "to_uppercase") { output_string(pg, ssk.to_uppercase(s)); return true; }
        tbh::FheString t;
        if (clear) t = ssk.trivial_string(cl);
        else { if (!need(2)) return false; t = ssk.input_string(a[1]); }
        if (f == "eq") pg.output(ssk.eq(s, t));
        else if (f == "ne") pg.output(ssk.ne(s, t));
        else if (f == "lt") pg.output(ssk.lt(s, t));
        else if (f == "le") pg.output(ssk.le(s, t));
        else if (f == "gt") pg.output(ssk.gt(s, t));
        else if (f == "ge") pg.output(ssk.ge(s, t));
        else if (f == "eq_ignore_case") pg.output(ssk.eq_ignore_case(s, t));
        else if (f == "contains") pg.output(ssk.contains(s, t));
        else if (f == "starts_with") pg.output(ssk.starts_with(s, t));
        else if (f == "ends_with") pg.output(ssk.ends_with(s, t));
        else if (f == "find") { auto r = ssk.find(s, t); pg.output(r.first); output_radix(pg, r.second); }
        else if (f == "rfind") { auto r = ssk.rfind(s, t); pg.output(r.first); output_radix(pg, r.second); }
        else if (f == "find_windows") {           // multi-GPU shard: first match among windows [a[2], a[3]) as a global index
            if (!need(4)) return false;
            if (a[2] > a[3]) { err = "find_windows: bad window range"; return false; }
            auto r = ssk.find_range(s, t, a[2], a[3]); pg.output(r.first); output_radix(pg, r.second);
        }
        else if (f == "cmp_sign") pg.output(ssk.compare_sign(s, t));     // multi-GPU shard of lt / le / gt / ge: the range's sign block
        else if (f == "contains_windows") {
            // multi-GPU shard: match flag OR-reduced over windows [a[2], a[3]) only -> one boolean block
            if (!need(4)) return false;
            const size_t nwin = t.len() <= s.len() ? s.len() - t.len() + 1 : 0;
            if (a[3] > nwin || a[2] >= a[3]) { err = "contains_windows: bad window range"; return false; }
            std::vector<tbh::Ct> mine = ssk.window_matches(s, t, a[2], a[3]);
            pg.output(isk.is_at_least_one_comparisons_block_true(mine));
        }
        else { err = "unknown string op: " + op; return false; }
        return true;
    }
    err = "unknown op: " + op;
    return false;
}

// frees the device copy of a program on the device it was uploaded to (no context needed: the owner may already be gone)
void release_device_state(tfhe_b200_program &h) {
    if (h.device < 0) return;
    DeviceGuard g(h.device);
    for (DevBuf *b : {&h.d_luts, &h.d_lin, &h.d_terms, &h.d_pbs_in, &h.d_pbs_out, &h.d_pbs_lut, &h.d_arena, &h.d_out_slots, &h.d_out_rows}) b->release();
    if (h.ev0) cudaEventDestroy(h.ev0);
    if (h.ev1) cudaEventDestroy(h.ev1);
    h.ev0 = h.ev1 = nullptr;
    h.owner_id = 0;
    h.device = -1;
}

bool same_params(const tfhe_b200_params &q, const tbh::Params &p) {
    return q.lwe_dim == p.lwe_dim && q.glwe_dim == p.glwe_dim && q.poly_size == p.poly_size && q.pbs_base_log == p.pbs_base_log &&
           q.pbs_level == p.pbs_level && q.ks_base_log == p.ks_base_log && q.ks_level == p.ks_level &&
           q.grouping_factor == p.grouping_factor && q.msg_mod == p.msg_mod && q.carry_mod == p.carry_mod;
}

// first run on this context: upload the program (accumulators, instruction arrays) and size its arena.  A program bound to another
// context is re-bound (its old device copy is dropped); `owner_id` is only set once every upload has succeeded.
int bind_program(tfhe_b200_ctx *c, tfhe_b200_program *h) {
    if (h->owner_id == c->id) return 0;
    release_device_state(*h);
    const tbh::Program &pg = *h->prog;
    cudaStream_t s = c->stream;
    const size_t L = c->big_len();
    h->device = c->device;            // from here on release_device_state() knows where the partial uploads live
    TB_CUDA(cudaEventCreate(&h->ev0));
    TB_CUDA(cudaEventCreate(&h->ev1));
    std::vector<uint64_t> accs(std::max<size_t>(h->luts->size(), 1) * h->p.lut_len());
    for (size_t u = 0; u < h->luts->size(); ++u) h->luts->fill_accumulator(uint32_t(u), accs.data() + u * h->p.lut_len());
    TB_CUDA(h->d_luts.reserve(accs.size() * 8));
    TB_CUDA(cudaMemcpyAsync(h->d_luts.p, accs.data(), accs.size() * 8, cudaMemcpyHostToDevice, s));
    std::vector<tbk::LinInstr> lin;
    std::vector<uint32_t> pin, pout, plut;
    for (auto &l : pg.levels()) {
        for (auto &i : l.lin) lin.push_back({i.out_slot, i.term_begin, i.term_end, 0u, i.body_add});
        for (auto &j : l.pbs) { pin.push_back(j.in_slot); pout.push_back(j.out_slot); plut.push_back(j.lut); }
    }
    std::vector<tbk::LinTerm> terms;
    for (auto &t : pg.terms()) terms.push_back({t.slot, 0u, t.coef});
    auto up = [&](DevBuf &b, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = b.reserve(std::max<size_t>(bytes, 16));
        if (e != cudaSuccess || bytes == 0) return e;
        return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s);
    };
    TB_CUDA(up(h->d_lin, lin.data(), lin.size() * sizeof(tbk::LinInstr)));
    TB_CUDA(up(h->d_terms, terms.data(), terms.size() * sizeof(tbk::LinTerm)));
    TB_CUDA(up(h->d_pbs_in, pin.data(), pin.size() * 4));
    TB_CUDA(up(h->d_pbs_out, pout.data(), pout.size() * 4));
    TB_CUDA(up(h->d_pbs_lut, plut.data(), plut.size() * 4));
    TB_CUDA(up(h->d_out_slots, pg.outputs().data(), pg.outputs().size() * 4));
    TB_CUDA(h->d_out_rows.reserve(std::max<size_t>(pg.outputs().size(), 1) * L * 8));
    TB_CUDA(h->d_arena.reserve(std::max<size_t>(pg.n_slots(), 1) * L * 8));
    TB_CUDA(cudaStreamSynchronize(s));   // the staging vectors above go out of scope
    h->owner_id = c->id;
    return 0;
}

// enqueues every level of the program on stream `s`: inputs are already in arena rows [0, n_inputs); the output rows end up in `d_out`
int enqueue_levels(tfhe_b200_ctx *c, tfhe_b200_program *h, uint64_t *d_out, cudaStream_t s) {
    const tbh::Program &pg = *h->prog;
    const size_t L = c->big_len();
    uint64_t *arena = (uint64_t *)h->d_arena.p;
    TB_CUDA(cudaEventRecord(h->ev0, s));
    size_t max_w = 0;
    for (auto &l : pg.levels()) max_w = std::max(max_w, l.pbs.size());
    TB_CUDA(c->d_small.reserve_on(std::max<size_t>(max_w, 1) * c->small_len() * 8, s));
    for (size_t lv = 0; lv < pg.levels().size(); ++lv) {
        const uint32_t l0 = h->lin_off[lv], l1 = h->lin_off[lv + 1], p0 = h->pbs_off[lv], p1 = h->pbs_off[lv + 1];
        if (l1 > l0) {
            TB_CUDA(tbk::launch_linear(arena, (const tbk::LinInstr *)h->d_lin.p + l0, (const tbk::LinTerm *)h->d_terms.p, int(l1 - l0), int(L), s));
            c->launches += 1;
        }
        if (p1 > p0) {
            const size_t w = p1 - p0;
            const bool fused = tbc::fused_supported(c);
            if (tbc::do_keyswitch(c, arena, (uint64_t *)c->d_small.p, w, s, (const uint32_t *)h->d_pbs_in.p + p0, nullptr, fused)) return 1;
            if (tbc::do_pbs(c, (const uint64_t *)c->d_small.p, (const uint32_t *)h->d_pbs_lut.p + p0, (const uint64_t *)h->d_luts.p, arena, w,
                            c->p.lwe_dim, s, (const uint32_t *)h->d_pbs_out.p + p0, fused))
                return 1;
        }
    }
    TB_CUDA(cudaEventRecord(h->ev1, s));
    if (!pg.outputs().empty()) {   // gather the output rows on the device
        TB_CUDA(tbk::launch_gather_rows(arena, (const uint32_t *)h->d_out_slots.p, d_out, (int)pg.outputs().size(), (int)L, s));
        c->launches += 1;
    }
    return 0;
}

}  // namespace

extern "C" {

int tfhe_b200_program_build(const tfhe_b200_params *params, const char *op, const uint64_t *args, size_t n_args,
                            const char *clear_operand, tfhe_b200_program **out) {
    if (!out) return fail("null out pointer");
    *out = nullptr;
    if (!params || !op) return fail("null argument");
    if (params->msg_mod < 2 || params->carry_mod < 1 || (params->poly_size % (params->msg_mod * params->carry_mod)) != 0)
        return fail("bad message / carry modulus");
    auto h = std::make_unique<tfhe_b200_program>();
    h->p = to_host_params(*params);
    h->luts = std::make_unique<tbh::LutRegistry>(h->p);
    h->prog = std::make_unique<tbh::Program>(h->p, *h->luts);
    h->op = op;
    std::string err;
    try {
        if (!record(*h, op, args, n_args, clear_operand, err)) return fail(err);
    } catch (const std::exception &e) {
        return fail(std::string(op) + ": " + e.what());
    }
    uint32_t lo = 0, po = 0;
    for (auto &l : h->prog->levels()) {
        h->lin_off.push_back(lo); h->pbs_off.push_back(po);
        lo += uint32_t(l.lin.size()); po += uint32_t(l.pbs.size());
    }
    h->lin_off.push_back(lo); h->pbs_off.push_back(po);
    *out = h.release();
    return 0;
}

int tfhe_b200_program_destroy(tfhe_b200_program *h) {
    if (!h) return 0;
    release_device_state(*h);
    delete h;
    return 0;
}

/* counts[8] = {n_inputs, n_outputs, n_slots, n_levels, n_lin, n_terms, n_pbs, n_luts}; counts[8] = trivial PBS folded on the host */
int tfhe_b200_program_counts(const tfhe_b200_program *h, uint64_t counts[9]) {
    if (!h || !counts) return fail("null argument");
    const tbh::Program &pg = *h->prog;
    counts[0] = pg.n_inputs(); counts[1] = pg.outputs().size(); counts[2] = pg.n_slots(); counts[3] = pg.levels().size();
    counts[4] = h->lin_off.back(); counts[5] = pg.terms().size(); counts[6] = h->pbs_off.back(); counts[7] = h->luts->size();
    counts[8] = pg.n_trivial_pbs();
    return 0;
}

/* Copies the recorded program out (for inspection and for executing it with another backend, e.g. the CPU oracle in tests).
 * Array sizes follow tfhe_b200_program_counts.  Any pointer may be NULL to skip that part. */
int tfhe_b200_program_copy(const tfhe_b200_program *h, uint32_t *level_lin_off, uint32_t *level_pbs_off, uint32_t *lin_out_tb_te,
                           uint64_t *lin_body, uint32_t *term_slot, int64_t *term_coef, uint32_t *pbs_in_out_lut,
                           uint64_t *lut_tables, uint64_t *lut_degrees, uint32_t *outputs, uint64_t *output_degree_noise) {
    if (!h) return fail("null argument");
    const tbh::Program &pg = *h->prog;
    if (level_lin_off) std::copy(h->lin_off.begin(), h->lin_off.end(), level_lin_off);
    if (level_pbs_off) std::copy(h->pbs_off.begin(), h->pbs_off.end(), level_pbs_off);
    size_t li = 0, pi = 0;
    for (auto &l : pg.levels()) {
        for (auto &i : l.lin) {
            if (lin_out_tb_te) { lin_out_tb_te[3 * li] = i.out_slot; lin_out_tb_te[3 * li + 1] = i.term_begin; lin_out_tb_te[3 * li + 2] = i.term_end; }
            if (lin_body) lin_body[li] = i.body_add;
            ++li;
        }
        for (auto &j : l.pbs) {
            if (pbs_in_out_lut) { pbs_in_out_lut[3 * pi] = j.in_slot; pbs_in_out_lut[3 * pi + 1] = j.out_slot; pbs_in_out_lut[3 * pi + 2] = j.lut; }
            ++pi;
        }
    }
    for (size_t t = 0; t < pg.terms().size(); ++t) {
        if (term_slot) term_slot[t] = pg.terms()[t].slot;
        if (term_coef) term_coef[t] = pg.terms()[t].coef;
    }
    for (size_t u = 0; u < h->luts->size(); ++u) {
        if (lut_tables) std::copy(h->luts->lut(u).table.begin(), h->luts->lut(u).table.end(), lut_tables + u * h->p.total_mod());
        if (lut_degrees) lut_degrees[u] = h->luts->lut(u).degree;
    }
    for (size_t o = 0; o < pg.outputs().size(); ++o) {
        if (outputs) outputs[o] = pg.outputs()[o];
        if (output_degree_noise) { output_degree_noise[2 * o] = pg.output_meta()[o].degree; output_degree_noise[2 * o + 1] = pg.output_meta()[o].noise; }
    }
    return 0;
}

/* Accumulators of the program's LUTs as ServerKey::generate_lookup_table would build them (n_luts x (k+1)*N words). */
int tfhe_b200_program_accumulators(const tfhe_b200_program *h, uint64_t *accs) {
    if (!h || !accs) return fail("null argument");
    for (size_t u = 0; u < h->luts->size(); ++u) h->luts->fill_accumulator(uint32_t(u), accs + u * h->p.lut_len());
    return 0;
}

/* Executes the program on `ctx` (its keys must be uploaded).  inputs: n_inputs x (k*N+1) words (host), in the order the
 * operation declares its operands; outputs: n_outputs x (k*N+1) words (host). */
int tfhe_b200_program_run(tfhe_b200_ctx *c, tfhe_b200_program *h, const uint64_t *inputs, uint64_t *outputs) {
    if (!c || !h) return fail("null argument");
    const tbh::Program &pg = *h->prog;
    if ((pg.n_inputs() && !inputs) || (!pg.outputs().empty() && !outputs)) return fail("null buffer");
    if (!same_params(c->p, h->p)) return fail("program was recorded for other parameters than the context");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaStream_t s = c->stream;
    const size_t L = c->big_len();
    if (bind_program(c, h)) return 1;
    if (pg.n_inputs()) TB_CUDA(cudaMemcpyAsync(h->d_arena.p, inputs, (size_t)pg.n_inputs() * L * 8, cudaMemcpyHostToDevice, s));
    if (enqueue_levels(c, h, (uint64_t *)h->d_out_rows.p, s)) return 1;
    if (!pg.outputs().empty())     // one copy back
        TB_CUDA(cudaMemcpyAsync(outputs, h->d_out_rows.p, pg.outputs().size() * L * 8, cudaMemcpyDeviceToHost, s));
    TB_CUDA(cudaStreamSynchronize(s));
    TB_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return 0;
}

/* The same on DEVICE buffers of the context's GPU, enqueued on `cuda_stream` (NULL = the context's own stream) WITHOUT synchronising:
 * d_inputs n_inputs x (k*N+1) words, d_outputs n_outputs x (k*N+1) words.  This is what chains a rank's share of a sharded string
 * operation, the exchange of the narrow-end blocks and the finishing program without a host round trip (multi_gpu.py).
 * One stream at a time per context: the keyswitch scratch and the program's arena are per context / per program. */
int tfhe_b200_program_run_device(tfhe_b200_ctx *c, tfhe_b200_program *h, const uint64_t *d_inputs, uint64_t *d_outputs, void *cuda_stream) {
    if (!c || !h) return fail("null argument");
    const tbh::Program &pg = *h->prog;
    if ((pg.n_inputs() && !d_inputs) || (!pg.outputs().empty() && !d_outputs)) return fail("null buffer");
    if (!same_params(c->p, h->p)) return fail("program was recorded for other parameters than the context");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    const size_t L = c->big_len();
    if (bind_program(c, h)) return 1;
    if (pg.n_inputs()) TB_CUDA(cudaMemcpyAsync(h->d_arena.p, d_inputs, (size_t)pg.n_inputs() * L * 8, cudaMemcpyDeviceToDevice, s));
    h->last_ms = -1.f;    // not synchronised: tfhe_b200_program_last_ms reads the events on demand
    return enqueue_levels(c, h, d_outputs, s);
}

/* Device time (ms, CUDA events on the context stream) of the kernels of the last tfhe_b200_program_run. */
int tfhe_b200_program_last_ms(const tfhe_b200_program *h, float *ms) {
    if (!h || !ms) return fail("null argument");
    if (h->last_ms < 0.f && h->ev1) {   // last run was tfhe_b200_program_run_device: wait for its end event
        DeviceGuard g(h->device);
        TB_CUDA(cudaEventSynchronize(h->ev1));
        TB_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
        return 0;
    }
    *ms = h->last_ms;
    return 0;
}

}  // extern "C"
