/* c_abi_smoke.c — proves include/lorikeet_b200.h is valid C11 (compiled with `gcc -std=c11 -Wall -Wextra -pedantic`),
 * and prints the size and field offsets of every struct that crosses the ABI as JSON, which tests/test_bindings.py compares
 * with the Python ctypes mirror (lorikeet_b200/_lib.py).  Links nothing: no entry point is called. */
#include <stddef.h>
#include <stdio.h>

#include "lorikeet_b200.h"
#include "lorikeet_b200_synth.h"

#define FIELD(T, f) printf("%s[\"%s\", %zu]", first ? "" : ", ", #f, offsetof(T, f)), first = 0
#define BEGIN(T) printf("%s\"%s\": {\"size\": %zu, \"fields\": [", comma ? ", " : "", #T, sizeof(T)), comma = 1, first = 1
#define END() printf("]}")

int main(void) {
    int comma = 0, first = 1;
    printf("{");
    BEGIN(lk_sw_params);
    FIELD(lk_sw_params, match_value); FIELD(lk_sw_params, mismatch_penalty); FIELD(lk_sw_params, gap_open_penalty);
    FIELD(lk_sw_params, gap_extend_penalty);
    END();
    BEGIN(lk_pairhmm_batch);
    FIELD(lk_pairhmm_batch, n_regions); FIELD(lk_pairhmm_batch, region_read_start); FIELD(lk_pairhmm_batch, region_hap_start);
    FIELD(lk_pairhmm_batch, read_off); FIELD(lk_pairhmm_batch, read_bases); FIELD(lk_pairhmm_batch, read_qual);
    FIELD(lk_pairhmm_batch, read_ins); FIELD(lk_pairhmm_batch, read_del); FIELD(lk_pairhmm_batch, read_gcp);
    FIELD(lk_pairhmm_batch, gcp_constant); FIELD(lk_pairhmm_batch, hap_off); FIELD(lk_pairhmm_batch, hap_bases);
    FIELD(lk_pairhmm_batch, region_sample_start); FIELD(lk_pairhmm_batch, sample_read_start);
    END();
    BEGIN(lk_realign_batch);
    FIELD(lk_realign_batch, n_regions); FIELD(lk_realign_batch, region_read_start); FIELD(lk_realign_batch, region_hap_start);
    FIELD(lk_realign_batch, values); FIELD(lk_realign_batch, region_n_evidence); FIELD(lk_realign_batch, ref_allele_index);
    FIELD(lk_realign_batch, reference_start); FIELD(lk_realign_batch, hap_off); FIELD(lk_realign_batch, hap_bases);
    FIELD(lk_realign_batch, hap_cigar_off); FIELD(lk_realign_batch, hap_cigars); FIELD(lk_realign_batch, hap_start_wrt_ref);
    FIELD(lk_realign_batch, read_off); FIELD(lk_realign_batch, read_bases); FIELD(lk_realign_batch, read_cigar_off);
    FIELD(lk_realign_batch, read_cigars); FIELD(lk_realign_batch, read_orig_len); FIELD(lk_realign_batch, informative_threshold);
    FIELD(lk_realign_batch, region_sample_start); FIELD(lk_realign_batch, sample_read_start); FIELD(lk_realign_batch, sample_n_evidence);
    END();
    BEGIN(lk_engine_args);
    FIELD(lk_engine_args, constant_gcp); FIELD(lk_engine_args, log10_global_read_mismapping_rate); FIELD(lk_engine_args, pcr_error_model);
    FIELD(lk_engine_args, base_quality_score_threshold); FIELD(lk_engine_args, dynamic_read_disqualification);
    FIELD(lk_engine_args, read_disqualification_scale); FIELD(lk_engine_args, expected_error_rate_per_base);
    FIELD(lk_engine_args, symmetrically_normalize_alleles_to_reference); FIELD(lk_engine_args, disable_cap_read_qualities_to_mapq);
    FIELD(lk_engine_args, modify_soft_clipped_bases);
    END();
    BEGIN(lk_synth_params);
    FIELD(lk_synth_params, seed_base); FIELD(lk_synth_params, n_reads); FIELD(lk_synth_params, read_len); FIELD(lk_synth_params, n_haps);
    FIELD(lk_synth_params, hap_len); FIELD(lk_synth_params, min_edits); FIELD(lk_synth_params, max_edits);
    FIELD(lk_synth_params, hap_edit_rate); FIELD(lk_synth_params, qual_profile); FIELD(lk_synth_params, read_indel_rate);
    FIELD(lk_synth_params, pcr_model); FIELD(lk_synth_params, base_quality_score_threshold); FIELD(lk_synth_params, mapq);
    END();
    printf(", \"enums\": {\"LK_OK\": %d, \"LK_ERR_INVALID\": %d, \"LK_ERR_CUDA\": %d, \"LK_ERR_NOMEM\": %d, \"LK_ERR_CAPACITY\": %d, "
           "\"LK_SW_SOFTCLIP\": %d, \"LK_SW_INDEL\": %d, \"LK_SW_LEADING_INDEL\": %d, \"LK_SW_IGNORE\": %d, \"LK_PCR_NONE\": %d, "
           "\"LK_PCR_HOSTILE\": %d, \"LK_PCR_AGGRESSIVE\": %d, \"LK_PCR_CONSERVATIVE\": %d, \"LK_FLAG_RESCUED\": %d, \"LK_FLAG_NEG_INF\": %d}",
           LK_OK, LK_ERR_INVALID, LK_ERR_CUDA, LK_ERR_NOMEM, LK_ERR_CAPACITY, LK_SW_SOFTCLIP, LK_SW_INDEL, LK_SW_LEADING_INDEL,
           LK_SW_IGNORE, LK_PCR_NONE, LK_PCR_HOSTILE, LK_PCR_AGGRESSIVE, LK_PCR_CONSERVATIVE, LK_FLAG_RESCUED, LK_FLAG_NEG_INF);
    printf("}\n");
    return 0;
}
