/*
 * hp_oracle.h -- C entry points of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This directory is a from-scratch CPU restatement of the reference algorithm
 * (PacificBiosciences/HiPhase v1.5.0, src/astar_phaser.rs, src/wfa_graph.rs, src/data_types/read_segments.rs,
 * src/read_parsing.rs:121-503 and 790-851, src/sequence_alignment.rs).  It exists to CHECK the CUDA path and to be timed as the CPU baseline.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Nothing under hiphase_b200/ links, imports or calls it.
 *
 * Parity pinning:
 *   - ReadSegment (clip / score / partial score / collapse), AstarNode costs+priorities, PQueueHapTracker and all
 *     19 WFA tests of the reference, plus edit_distance (sequence_alignment.rs:45-76) and match_allele /
 *     closest_allele (variants.rs:668-846), are transcribed as data in tests/golden/ and checked against this oracle.
 *   - astar_solver / astar_subsolver / calculate_astar_heuristic have NO known-answer test in the reference
 *     (SURVEY.md section 8c): for those three functions parity is UNPINNED -- the oracle follows the code text of
 *     src/astar_phaser.rs:246-633 and is cross-checked against an independent pure-Python restatement
 *     (tests/pyref.py), brute-force MEC optima on tiny blocks, and the reference's own asserts.
 *   - The Rust reference cannot be compiled in this image (no cargo/rustc, no vendored crates, no htslib), so
 *     there is no oracle/_ref build.
 *
 * The structs are the ones of include/hiphase_b200.h so the same packed batch feeds the oracle and the product.
 */
#ifndef HP_ORACLE_H
#define HP_ORACLE_H

#include "../include/hiphase_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ReadSegment (src/data_types/read_segments.rs) ---- */
/* ReadSegment::new: clip to [first set, last set+1).  Returns region in *start,*end. (read_segments.rs:40-62) */
void hpo_read_segment_region(const uint8_t* alleles, uint64_t n, uint64_t* start, uint64_t* end);
/* score_partial_haplotype (read_segments.rs:177-206) of a clipped read against hap[0..hap_len) at offset. */
uint64_t hpo_score_partial(uint64_t start, uint64_t end, const uint8_t* alleles, const uint8_t* quals,
                           const uint8_t* hap, uint64_t hap_len, uint64_t offset);
/* ReadSegment::collapse (read_segments.rs:71-121): k full-length (n cells each) mappings -> one row; returns 0 ok. */
int hpo_collapse(uint32_t k, uint64_t n, const uint8_t* alleles /*k*n*/, const uint8_t* quals /*k*n*/,
                 uint8_t* out_alleles /*n, NoOverlap outside region*/, uint8_t* out_quals, uint64_t* start, uint64_t* end);

/* matrix assembly (read_segments.rs:40-121, 151-155; read_parsing.rs:612-629); returns 0 ok */
int hpo_assemble_blocks(const hp_rows_batch* rows, hp_assembled* out);

/* ---- AstarNode / tracker probes for the reference's unit tests (astar_phaser.rs:663-798) ---- */
/*
 * Walks one path of a single block: child d appends (a1[d], a2[d]) with heuristic H[d+1] (main-solver
 * convention, hap_offset 0).  Writes per depth (1..n_steps) frozen, total, num_hets.  Returns 0 ok.
 */
int hpo_astar_node_path(const hp_block_batch* one_block, const uint64_t* H, uint32_t n_steps,
                        const uint8_t* a1, const uint8_t* a2,
                        uint64_t* frozen, uint64_t* total, uint64_t* hets);
/* PQueueHapTracker script: ops[i] = 0 add / 1 remove / 2 increase_threshold, value in vals[i]; lens[i] = len() after op. */
int hpo_tracker_script(uint32_t max_len, uint32_t n_ops, const uint8_t* ops, const uint32_t* vals, uint64_t* lens);

/* ---- astar_solver (astar_phaser.rs:426-633) over a batch; threads = worker pool size (src/main.rs:332-408) ---- */
int hpo_astar_solve_batch(const hp_params* params, const hp_block_batch* batch, hp_astar_out* out, int threads);
/* one sub-problem: astar_subsolver (astar_phaser.rs:311-405) on block 0 of the batch; H has n_var+1 entries. */
int hpo_astar_subsolver(const hp_params* params, const hp_block_batch* one_block, uint64_t problem_offset,
                        uint64_t problem_size, const uint64_t* H, uint64_t* max_cost, uint64_t* solved);

/* ---- post-solve (src/phaser.rs:350-388, 546-569, 714-750) ---- */
int hpo_post_solve_batch(const hp_block_batch* batch, const int64_t* var_pos, const uint8_t* h1, const uint8_t* h2, hp_post_out* out);

/* ---- local realignment (src/read_parsing.rs:121-503, src/sequence_alignment.rs:6-38, variants.rs:598-641) ---- */
uint64_t hpo_edit_distance(const uint8_t* a, uint64_t la, const uint8_t* b, uint64_t lb);
uint8_t  hpo_match_allele(const hp_local_batch* batch, uint32_t variant, const uint8_t* seq, uint64_t n);
uint8_t  hpo_closest_allele_clip(const hp_local_batch* batch, uint32_t variant, const uint8_t* seq, uint64_t n,
                                 uint64_t head_clip, uint64_t tail_clip, uint64_t* d0, uint64_t* d1);
int      hpo_local_realign_batch(const hp_local_batch* batch, hp_local_out* out);

/* ---- WFA graph (src/wfa_graph.rs) ---- */
typedef struct hpo_graph hpo_graph;
hpo_graph* hpo_graph_new(uint64_t max_edit_distance);
void       hpo_graph_free(hpo_graph* g);
/* add_node (wfa_graph.rs:298-331): returns the node index, or -1 on the reference's bail! conditions. */
int64_t    hpo_graph_add_node(hpo_graph* g, const uint8_t* seq, uint64_t len, const uint64_t* parents, uint32_t n_parents);
uint64_t   hpo_graph_num_nodes(const hpo_graph* g);
/* from_reference_variants_with_hom (wfa_graph.rs:119-284) for job j of the batch.  Returns NULL on error. */
hpo_graph* hpo_graph_from_job(const hp_wfa_batch* batch, uint32_t job, uint64_t max_edit_distance);
/* NodeAlleleMap entries of a graph built by hpo_graph_from_job: triplets (node, var_index, allele); returns count. */
uint64_t   hpo_graph_allele_map(const hpo_graph* g, uint64_t* triplets, uint64_t cap);
/* flatten (for feeding hp_wfa_graph_align): sizes then arrays. */
void       hpo_graph_sizes(const hpo_graph* g, uint64_t* n_nodes, uint64_t* n_seq, uint64_t* n_parents);
void       hpo_graph_flatten(const hpo_graph* g, uint8_t* seq, uint64_t* seq_off, uint32_t* parent_idx, uint64_t* parent_off);
/*
 * edit_distance_with_pruning (wfa_graph.rs:350-650).  prune_distance UINT64_MAX = disabled.
 * Returns HP_WFA_OK / HP_WFA_MAX_EDIT_DISTANCE; writes score, traversed node ids (sorted) and their count.
 * shuffle_seed != 0 visits the diagonals of each node in a seeded random order (order-independence check).
 */
int hpo_graph_edit_distance(const hpo_graph* g, const uint8_t* read, uint64_t read_len, uint64_t prune_distance,
                            uint64_t* score, uint64_t* traversed, uint64_t* n_traversed,
                            hp_wfa_counters* counters, uint64_t shuffle_seed);

/* whole batch: graph build + alignment + allele/qual rows (read_parsing.rs:769-851). */
int hpo_wfa_align_batch(const hp_params* params, const hp_wfa_batch* batch, hp_wfa_out* out, int threads);

/* ---- the read loop of load_full_read_segments (src/read_parsing.rs:545-629), one mapping at a time in BAM order ---- */
int hpo_realign_block_batch(const hp_params* params, const hp_realign_batch* batch, hp_realign_out* out);
/* the pre-WFA half of global_realignment (src/read_parsing.rs:672-742) */
int hpo_wfa_plan_batch(const hp_plan_batch* batch, hp_plan_out* out);

#ifdef __cplusplus
}
#endif
#endif
