/* oracle.h -- declarations for the CPU oracle (TEST INFRASTRUCTURE ONLY; see oracle.c). */
#ifndef SOURMASH_B200_ORACLE_H
#define SOURMASH_B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_mh orc_mh;

uint64_t orc_hash_murmur(const uint8_t *data, size_t len, uint64_t seed);
uint64_t orc_max_hash_for_scaled(uint64_t scaled);
uint64_t orc_scaled_for_max_hash(uint64_t max_hash);
int64_t orc_seq_to_hashes(const uint8_t *seq, size_t len, uint32_t ksize, uint64_t seed,
                          int force, int keep_zeros, uint64_t *out, int64_t *err_index,
                          int64_t *n_before_err);

orc_mh *orc_mh_new(uint64_t scaled, uint32_t ksize, uint64_t seed, int track, uint32_t num);
void orc_mh_free(orc_mh *m);
size_t orc_mh_size(const orc_mh *m);
const uint64_t *orc_mh_mins(const orc_mh *m);
const uint64_t *orc_mh_abunds(const orc_mh *m);
uint64_t orc_mh_max_hash(const orc_mh *m);
void orc_mh_clear(orc_mh *m);
void orc_mh_remove_hash(orc_mh *m, uint64_t h);
void orc_mh_add_hash(orc_mh *m, uint64_t h);
void orc_mh_add_hash_with_abundance(orc_mh *m, uint64_t h, uint64_t abundance);
void orc_mh_add_many(orc_mh *m, const uint64_t *h, size_t n);
int64_t orc_mh_add_sequence(orc_mh *m, const uint8_t *seq, size_t len, int force);
void orc_mh_merge(orc_mh *m, const orc_mh *o);
void orc_md5sum(uint32_t ksize, const uint64_t *mins, size_t n, char out_hex[33]);

uint64_t orc_count_common(const uint64_t *a, size_t na, const uint64_t *b, size_t nb);
void orc_intersection_size(const uint64_t *a, size_t na, const uint64_t *b, size_t nb,
                           uint64_t *common, uint64_t *union_size);
void orc_intersection_size_num(const uint64_t *a, size_t na, const uint64_t *b, size_t nb,
                               uint32_t num, uint64_t *common, uint64_t *union_size);
double orc_jaccard(const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint32_t num);
double orc_angular_similarity(const uint64_t *a, const uint64_t *aa, size_t na,
                              const uint64_t *b, const uint64_t *ba, size_t nb);
size_t orc_downsample_count(const uint64_t *a, size_t na, uint64_t new_max_hash);

void orc_compare_all_pairs(const uint64_t *hashes, const uint64_t *offsets, size_t n,
                           uint32_t num, size_t first_row, size_t n_rows, double *out,
                           int nthreads);
void orc_pairwise_common(const uint64_t *hashes, const uint64_t *offsets, size_t n,
                         size_t first_row, size_t n_rows, uint32_t *out, int nthreads);
void orc_one_vs_many(const uint64_t *q, size_t nq, const uint64_t *hashes,
                     const uint64_t *offsets, size_t n, uint64_t *common, int nthreads);
void orc_one_vs_many_bsearch(const uint64_t *q, size_t nq, const uint64_t *hashes,
                             const uint64_t *offsets, size_t n, uint64_t *common, int nthreads);
size_t orc_sketch_scaled(const uint8_t *seq, size_t len, uint32_t k, uint64_t seed,
                         uint64_t max_hash, uint64_t *dst, size_t dst_cap, uint64_t *n_kmers);
void orc_sketch_batch(const uint8_t *seqs, const uint64_t *seq_off, size_t n_seqs, uint32_t k,
                      uint64_t seed, uint64_t max_hash, uint64_t *out, const uint64_t *out_off,
                      uint64_t *out_n, int nthreads);

/* protein-family hashing (hash_function: 2 protein, 3 dayhoff, 4 hp; ksize = 3 x residues) */
uint8_t orc_translate_codon(const uint8_t *codon, size_t n);
uint8_t orc_aa_to_dayhoff(uint8_t aa);
uint8_t orc_aa_to_hp(uint8_t aa);
int64_t orc_seq_to_hashes_protein(const uint8_t *seq, size_t len, uint32_t ksize, uint64_t seed,
                                  int hash_function, int keep_zeros, uint64_t *out);
int64_t orc_seq_to_hashes_translate(const uint8_t *seq, size_t len, uint32_t ksize, uint64_t seed,
                                    int hash_function, int keep_zeros, uint64_t *out);
int64_t orc_mh_add_protein_family(orc_mh *m, const uint8_t *seq, size_t len, int hash_function,
                                  int input_is_protein);
#ifdef __cplusplus
}
#endif
#endif
