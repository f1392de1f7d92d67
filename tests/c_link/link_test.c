/* Plain-C consumer of include/tfhe_b200.h: proves that the header is self-sufficient C (no C++, no CUDA, no torch types) and that
 * libtfhe_b200.so resolves every entry point a host-language binding would declare.  Runs without a GPU: it only calls the entry
 * points that need none (version, program recording / inspection, wire parsing, error paths) and takes the address of the rest. */
#include <stdio.h>
#include <string.h>
#include "tfhe_b200.h"

typedef void (*any_fn)(void);
#define ADDR(f) do { any_fn p = (any_fn)(f); if (!p) { printf("missing %s\n", #f); return 2; } n_syms++; } while (0)

int main(void) {
    int n_syms = 0;
    ADDR(tfhe_b200_ctx_create); ADDR(tfhe_b200_ctx_destroy); ADDR(tfhe_b200_last_error); ADDR(tfhe_b200_set_ciphertext_modulus_log2);
    ADDR(tfhe_b200_upload_ksk); ADDR(tfhe_b200_upload_bsk_std); ADDR(tfhe_b200_upload_luts);
    ADDR(tfhe_b200_upload_seeded_ksk); ADDR(tfhe_b200_upload_seeded_bsk);
    ADDR(tfhe_b200_wire_parse_compressed_server_key); ADDR(tfhe_b200_load_compressed_server_key);
    ADDR(tfhe_b200_wire_read_ciphertexts); ADDR(tfhe_b200_wire_write_ciphertexts);
    ADDR(tfhe_b200_keyswitch_batch); ADDR(tfhe_b200_pbs_batch); ADDR(tfhe_b200_ks_pbs_batch); ADDR(tfhe_b200_pbs_ks_batch);
    ADDR(tfhe_b200_keyswitch_batch_device); ADDR(tfhe_b200_pbs_batch_device); ADDR(tfhe_b200_ks_pbs_batch_device);
    ADDR(tfhe_b200_synchronize); ADDR(tfhe_b200_pbs_batch_partial);
    ADDR(tfhe_b200_program_build); ADDR(tfhe_b200_program_destroy); ADDR(tfhe_b200_program_counts); ADDR(tfhe_b200_program_copy);
    ADDR(tfhe_b200_program_accumulators); ADDR(tfhe_b200_program_run); ADDR(tfhe_b200_program_run_device); ADDR(tfhe_b200_program_last_ms);
    ADDR(tfhe_b200_exchange_create); ADDR(tfhe_b200_exchange_handle); ADDR(tfhe_b200_exchange_attach); ADDR(tfhe_b200_exchange_attach_local);
    ADDR(tfhe_b200_exchange_send_rows); ADDR(tfhe_b200_exchange_gather_stride); ADDR(tfhe_b200_exchange_all_gather);
    ADDR(tfhe_b200_exchange_all_reduce_sum); ADDR(tfhe_b200_exchange_group_run); ADDR(tfhe_b200_exchange_destroy);
    ADDR(tfhe_b200_set_tuning); ADDR(tfhe_b200_kernel_launches); ADDR(tfhe_b200_time_last_kernels); ADDR(tfhe_b200_probe_fp64_tflops);
    ADDR(tfhe_b200_version); ADDR(tfhe_b200_plan_classic_level);

    printf("version: %s\n", tfhe_b200_version());
    size_t cut[3];
    tfhe_b200_plan_classic_level(1024, 148, 296, 1, cut);
    printf("level of 1024 ciphertexts: %zu + %zu + %zu\n", cut[0], cut[1], cut[2]);
    if (cut[0] + cut[1] + cut[2] != 1024) return 2;
    /* PARAM_MESSAGE_2_CARRY_2_KS_PBS, shortint/parameters/mod.rs:703-717 */
    tfhe_b200_params p = {742, 1, 2048, 23, 1, 3, 5, 0, 4, 4};
    tfhe_b200_program *prog = NULL;
    uint64_t args[2] = {8, 8};
    if (tfhe_b200_program_build(&p, "string_eq", args, 2, NULL, &prog) != 0 || !prog) { printf("program_build: %s\n", tfhe_b200_last_error()); return 3; }
    uint64_t counts[9];
    if (tfhe_b200_program_counts(prog, counts) != 0) return 4;
    printf("string_eq(8, 8): %llu inputs, %llu PBS in %llu levels\n", (unsigned long long)counts[0], (unsigned long long)counts[6], (unsigned long long)counts[3]);
    if (counts[0] != 64 || counts[6] != 36) return 5;
    tfhe_b200_program_destroy(prog);
    /* error conventions of tfhe/src/c_api/utils.rs:3-28: non-zero return, out-pointer nulled, message available */
    prog = (tfhe_b200_program *)1;
    if (tfhe_b200_program_build(&p, "no_such_op", args, 2, NULL, &prog) == 0 || prog != NULL) return 6;
    if (strlen(tfhe_b200_last_error()) == 0) return 7;
    tfhe_b200_wire_server_key view;
    const uint8_t junk[4] = {1, 2, 3, 4};
    if (tfhe_b200_wire_parse_compressed_server_key(junk, sizeof junk, &view) == 0) return 8;
    tfhe_b200_ctx *ctx = (tfhe_b200_ctx *)1;
    int rc = tfhe_b200_ctx_create(0, &p, &ctx);      /* fails without a GPU (no CPU fallback), succeeds on the GPU box */
    printf("ctx_create: rc = %d%s%s\n", rc, rc ? ", " : "", rc ? tfhe_b200_last_error() : "");
    if (rc != 0 && ctx != NULL) return 9;
    if (rc == 0) tfhe_b200_ctx_destroy(ctx);
    printf("ok: %d symbols resolved\n", n_syms);
    return 0;
}
