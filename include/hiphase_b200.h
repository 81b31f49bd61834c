/*
 * hiphase_b200.h -- C ABI of the B200-native per-block phasing hot path.
 *
 * This is the drop-in boundary.  The reference (PacificBiosciences/HiPhase v1.5.0) has no FFI of its own:
 * the hot path sits behind two internal Rust call sites, and each entry point below names the one it replaces.
 *
 *   call site 1  src/phaser.rs:541-543       astar_phaser::astar_solver(block, variants, read_segments,
 *                                            min_queue_size, queue_increment) -> AstarResult
 *                                            (src/astar_phaser.rs:426-633)
 *   call site 2  src/read_parsing.rs:769-780 WFAGraph::from_reference_variants_with_hom(...)   (src/wfa_graph.rs:119)
 *                                            WFAGraph::edit_distance_with_pruning(read, prune) (src/wfa_graph.rs:350)
 *                                            + traversed nodes -> allele/qual row              (src/read_parsing.rs:790-851)
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every buffer it passes in; the library never frees them.
 *   - every function returns an int status: HP_OK (0) or a negative HP_ERR_* code.  Conditions the reference
 *     answers with panic!/assert! (e.g. src/astar_phaser.rs:439, 631) come back as HP_ERR_INVALID_INPUT or as a
 *     per-block / per-job status, never as a crash.
 *   - "host" entry points take host pointers (pinned or pageable) and do the H2D/D2H copies themselves;
 *     "_device" entry points take device pointers on the context's GPU and a cudaStream_t (passed as void*).
 *   - there is NO CPU fallback: without a CUDA device hp_ctx_create fails with HP_ERR_NO_DEVICE.
 *   - re-entrancy: one hp_ctx per thread (the reference runs one solve_block per worker, src/main.rs:385-408);
 *     different contexts may be used concurrently.
 */
#ifndef HIPHASE_B200_H
#define HIPHASE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HP_ABI_VERSION 2

/* ---- status codes ---------------------------------------------------------------------------------------- */
#define HP_OK                      0
#define HP_ERR_INVALID_INPUT      -1   /* malformed batch (reference would panic / bail!)                      */
#define HP_ERR_NO_DEVICE          -2   /* no CUDA device / wrong architecture: the product never runs on CPU   */
#define HP_ERR_CUDA               -3   /* a CUDA runtime call failed; see hp_last_error()                      */
#define HP_ERR_UNSUPPORTED        -4   /* parameters outside what the kernels were built for                   */
#define HP_ERR_OUT_OF_MEMORY      -5
#define HP_ERR_INTERNAL           -6

/* per-block status written to hp_astar_out.status */
#define HP_BLOCK_OK                0
#define HP_BLOCK_IGNORED_NOT_NOOVERLAP  1  /* an ignored variant is not NoOverlap(3) in some read: astar_phaser.rs:435-442 */
#define HP_BLOCK_COST_OVERFLOW     2   /* sum of all quals in the block >= 2^31: 32-bit cost path refused       */
#define HP_BLOCK_QUEUE_OVERFLOW    3   /* internal: main queue outgrew its slab (retried transparently)         */
#define HP_BLOCK_ASSERT            4   /* a reference assert! would have fired (e.g. astar_phaser.rs:284, 529)  */
#define HP_BLOCK_TOO_DENSE         5   /* more than 65534 reads cover one variant: outside the kernel's range   */
#define HP_BLOCK_INDEX_EXHAUSTED   6   /* more than 2^32-16 nodes created: a larger slab cannot help (not retried) */

/* per-job status written to hp_wfa_out.status */
#define HP_WFA_OK                  0
#define HP_WFA_MAX_EDIT_DISTANCE   1   /* WFAGraphError::MaxEditDistance, src/wfa_graph.rs:645-648              */
#define HP_WFA_SKIPPED             2   /* job overlaps no het variant (read_parsing.rs:703-712)                 */
#define HP_WFA_WORKSPACE_OVERFLOW  3   /* internal: wave table outgrew its slab (retried transparently up to 4x) */
#define HP_WFA_GRAPH_TOO_LARGE     4   /* graph has more than 4096 nodes (> ~1300 variants in one read window):  */
                                       /* outside the kernel's range, NOT retried; the caller routes the read to */
                                       /* local realignment (hp_realign_block_batch does), as for MaxEditDistance */

/* AlleleType, src/data_types/read_segments.rs:5-16 */
#define HP_ALLELE_REFERENCE  0
#define HP_ALLELE_ALTERNATE  1
#define HP_ALLELE_AMBIGUOUS  2
#define HP_ALLELE_NOOVERLAP  3

/* VariantType, src/data_types/variants.rs:8-31 (numeric order of the Rust enum) */
#define HP_VT_SNV            0
#define HP_VT_INSERTION      1
#define HP_VT_DELETION       2
#define HP_VT_INDEL          3
#define HP_VT_SV_INSERTION   4
#define HP_VT_SV_DELETION    5
#define HP_VT_SV_DUPLICATION 6
#define HP_VT_SV_INVERSION   7
#define HP_VT_SV_BREAKEND    8
#define HP_VT_TANDEM_REPEAT  9
#define HP_VT_UNKNOWN       10

/* ---- parameters (defaults = src/cli.rs:186-226) ------------------------------------------------------------ */
typedef struct hp_params {
    uint32_t min_queue_size;        /* --phase-min-queue-size, default 1000 (cli.rs:215-219)                    */
    uint32_t queue_increment;       /* --phase-queue-increment, default 3   (cli.rs:222-226)                    */
    uint32_t wfa_prune_distance;    /* --global-pruning-distance, default 500; 0 = disabled (cli.rs:194-198,352)*/
    uint32_t wfa_max_edit_distance; /* --global-realignment-max-ed, default 500 (cli.rs:187-191)                */
} hp_params;

void hp_default_params(hp_params* p);

typedef struct hp_ctx hp_ctx;

int  hp_abi_version(void);
/* How this binary was built: ABI version, nvcc release, target architecture, compile date (bench.py records it). */
const char* hp_build_info(void);
/* device < 0 selects the current CUDA device. */
int  hp_ctx_create(const hp_params* params, int device, hp_ctx** out_ctx);
void hp_ctx_destroy(hp_ctx* ctx);
/* last error text of this context (or of the failed hp_ctx_create when ctx == NULL); never NULL. */
const char* hp_last_error(const hp_ctx* ctx);

/* ---- A* phasing: replaces astar_phaser::astar_solver (src/astar_phaser.rs:426) ---------------------------- */
/*
 * A batch of independent phase blocks in the reference's own u8 layout.  Block b owns
 *   variants   [var_off[b],  var_off[b+1])    -> ignored[], is_snv[], and the outputs h1[], h2[]
 *   reads      [read_off[b], read_off[b+1])   -> read_start[], read_end[], cell_off[]
 * Read r is one collapsed ReadSegment (src/data_types/read_segments.rs:19-62): region = [read_start[r], read_end[r])
 * in block-relative variant indices, and its clipped alleles / quals are the read_end-read_start bytes at
 * alleles[cell_off[r] ..], quals[cell_off[r] ..]   (cell_off[r+1]-cell_off[r] == read_end[r]-read_start[r]).
 */
typedef struct hp_block_batch {
    uint32_t        n_blocks;
    const uint64_t* var_off;     /* [n_blocks+1]                                                           */
    const uint64_t* read_off;    /* [n_blocks+1]                                                           */
    const uint32_t* read_start;  /* [n_reads]                                                              */
    const uint32_t* read_end;    /* [n_reads]                                                              */
    const uint64_t* cell_off;    /* [n_reads+1]                                                            */
    const uint8_t*  alleles;     /* [n_cells]   HP_ALLELE_*                                                */
    const uint8_t*  quals;       /* [n_cells]                                                              */
    const uint8_t*  ignored;     /* [n_vars]    Variant::is_ignored() (astar_phaser.rs:445-447)            */
    const uint8_t*  is_snv;      /* [n_vars]    Variant::get_type()==Snv (astar_phaser.rs:606)             */
} hp_block_batch;

/* PhaseStats as produced by astar_solver (astar_phaser.rs:599-621; writers/phase_stats.rs:159-173). */
typedef struct hp_phase_stats {
    uint64_t pruned_solutions;
    uint64_t estimated_cost;      /* H[0]                                                                  */
    uint64_t actual_cost;
    uint64_t phased_variants;
    uint64_t phased_snvs;
    uint64_t homozygous_variants;
    uint64_t skipped_variants;
} hp_phase_stats;

/* Exact work counters of the reference algorithm for one block (SURVEY.md section 8d "algorithmic bytes"). */
typedef struct hp_astar_counters {
    uint64_t evals;          /* AstarNode::new_extended_node calls (astar_phaser.rs:69), pre-pass + main     */
    uint64_t cells;          /* sum over evals and active reads of the window w_r one score_partial scans    */
    uint64_t sum_parent_len; /* sum over evals of the parent haplotype length                                */
    uint64_t pops;           /* queue pops, pre-pass + main                                                  */
} hp_astar_counters;

typedef struct hp_astar_out {
    uint8_t*           h1;         /* [n_vars]  haplotype_1, values 0/1, 2 for ignored variants               */
    uint8_t*           h2;         /* [n_vars]  haplotype_2                                                   */
    hp_phase_stats*    stats;      /* [n_blocks]                                                              */
    int32_t*           status;     /* [n_blocks] HP_BLOCK_*                                                   */
    uint64_t*          heuristic;  /* optional (may be NULL): [n_vars + n_blocks], block b at var_off[b]+b, N+1 entries */
    hp_astar_counters* counters;   /* optional (may be NULL): [n_blocks]                                      */
} hp_astar_out;

/* Host buffers in, host buffers out: H2D + kernels + D2H (the end-to-end call). */
int hp_astar_solve_batch(hp_ctx* ctx, const hp_block_batch* batch, hp_astar_out* out);

/*
 * Device-resident variant: every pointer inside *batch and *out is a device pointer on the context's GPU; the
 * work is enqueued on `stream` (a cudaStream_t) and the call returns without synchronising.  n_vars / n_reads /
 * n_cells are the array lengths and max_block_vars the largest block's variant count (sizes the queue records).
 * A block whose main queue outgrows its slab reports HP_BLOCK_QUEUE_OVERFLOW (hp_astar_solve_batch retries those
 * transparently with a larger slab; this entry point leaves the retry to the caller: check out->status[]).
 * Calls on different streams overlap on the device (each takes the next of the context's lanes, see hp_ctx_set_lanes).
 */
int hp_astar_solve_device(hp_ctx* ctx, const hp_block_batch* batch, uint64_t n_vars, uint64_t n_reads,
                          uint64_t n_cells, uint32_t max_block_vars, hp_astar_out* out, void* stream);

/* Single-block convenience with the shape of call site 1 (host buffers). */
int hp_astar_solve_one(hp_ctx* ctx, uint32_t n_var, uint32_t n_reads,
                       const uint32_t* read_start, const uint32_t* read_end, const uint64_t* cell_off,
                       const uint8_t* alleles, const uint8_t* quals,
                       const uint8_t* ignored, const uint8_t* is_snv,
                       uint8_t* h1, uint8_t* h2, hp_phase_stats* stats);


/*
 * Block service: the per-block call surface of the reference on top of batched launches.  HiPhase calls astar_solver from
 * up to --threads pool workers, one phase block each (src/main.rs:385-408, src/phaser.rs:541-543); a GPU wants many blocks per
 * launch.  A service owns one context and a dispatcher thread: hp_service_solve_one is thread-safe and BLOCKING with the
 * argument list of hp_astar_solve_one; blocks that arrive while the device is busy (or within linger_us of each other) are
 * packed into one batch and go out through hp_astar_submit on the context's lanes, so a long block of one batch does not hold
 * up the blocks that arrive after it.  One service per GPU, shared by all worker threads.
 */
typedef struct hp_service hp_service;
int  hp_service_create(const hp_params* params /* NULL = defaults */, int device, uint32_t max_batch_blocks /* 0 = 4096 */,
                       uint32_t linger_us /* how long the dispatcher waits for more blocks once one is pending; 0 = 200 */,
                       hp_service** out);
int  hp_service_solve_one(hp_service* svc, uint32_t n_var, uint32_t n_reads,
                          const uint32_t* read_start, const uint32_t* read_end, const uint64_t* cell_off,
                          const uint8_t* alleles, const uint8_t* quals, const uint8_t* ignored, const uint8_t* is_snv,
                          uint8_t* h1, uint8_t* h2, hp_phase_stats* stats);
/* batches launched / blocks solved so far (how well the calls were packed) */
int  hp_service_counters(const hp_service* svc, uint64_t* n_batches, uint64_t* n_blocks);
void hp_service_destroy(hp_service* svc);

/*
 * Streaming entry: the reference keeps 40 x threads jobs in flight and streams results back (src/main.rs:328, 344-355).
 * hp_astar_submit enqueues H2D + kernels + D2H of one batch on one of the context's lanes (own stream and workspaces;
 * hp_ctx_set_lanes, default 4) and returns without waiting; batches in flight on different lanes share the GPU, so the
 * long serial chain of one noisy block overlaps the next batches instead of idling the device.  The caller's buffers
 * (*batch arrays and *out arrays) must stay valid and untouched until hp_astar_wait returns.  Pinned buffers
 * (hp_host_alloc / hp_host_register) are copied from and to directly.  Pageable buffers (an ordinary Vec) are staged by the
 * library: submit gathers the inputs into the lane's pinned staging with a few host threads (it may then be reused at
 * once, though the contract above does not promise it), the results land in pinned staging and hp_astar_wait copies them
 * into *out -- so the results of a pageable job are in *out only after hp_astar_wait, not when hp_astar_poll reports done.
 * C3 end to end: 116 k blocks/s pinned, 104 k pageable (37 k before the staging: a cudaMemcpyAsync to pageable memory
 * blocks the caller until the stream gets there).
 * Which build of the solver kernels a batch gets is decided per launch: with other batches in flight and at least four
 * blocks per warp of the launch's share of the device it is the 20-warps-per-SM build (throughput), else the 16-warp one
 * (shortest chain per block).  Results do not depend on it.
 * hp_astar_wait blocks until the results are in *out (overflowed blocks are re-run there, as in hp_astar_solve_batch)
 * and releases the job.  Submitting more jobs than lanes waits for the oldest job's device work first.
 * HP_OK from wait / solve_batch means the call ran: each block's own outcome is in out->status[] (HP_BLOCK_*).
 */
typedef struct hp_astar_job hp_astar_job;
int hp_ctx_set_lanes(hp_ctx* ctx, int lanes);     /* 1..16 */
int hp_astar_submit(hp_ctx* ctx, const hp_block_batch* batch, hp_astar_out* out, hp_astar_job** job);
int hp_astar_poll(hp_ctx* ctx, hp_astar_job* job, int* done);   /* *done = 1 when hp_astar_wait would not block on the device */
int hp_astar_wait(hp_ctx* ctx, hp_astar_job* job);

/* Pinned host memory for the host entry points (a Rust Vec can be registered in place). */
int hp_host_alloc(void** ptr, size_t bytes);
int hp_host_free(void* ptr);
int hp_host_register(void* ptr, size_t bytes);
int hp_host_unregister(void* ptr);

/* ---- sharding phase blocks over the GPUs of one box --------------------------------------------------
 * Blocks share nothing (src/main.rs:385-408 runs them as independent pool jobs), so there is no data-path collective:
 * the work-queue hand-off is a cost-sorted deal of whole blocks before the solve, and one gather of the results after.
 */
/* Serial-chain cost model of a block: n_cells * min(n_var, 40) + n_var (the heuristic look-ahead is 40 variants). */
int hp_block_costs(uint64_t n_blocks, const uint32_t* n_var, const uint64_t* n_cells, uint64_t* cost);
/* Longest-processing-time-first deal: blocks by descending cost (ties: lower index first), each to the least loaded
 * shard (ties: lower shard).  shard_of[i] in [0, n_shards). */
int hp_lpt_partition(const uint64_t* cost, uint64_t n_blocks, uint32_t n_shards, uint32_t* shard_of);

/* NCCL communicator owned by the context (one process per GPU; NVLink 5 / NVSwitch).  Rank 0 calls hp_comm_unique_id and
 * hands the 128 bytes to every rank (any out-of-band channel); then every rank calls hp_comm_init. */
#define HP_COMM_ID_BYTES 128
int hp_comm_unique_id(uint8_t id[HP_COMM_ID_BYTES]);
int hp_comm_init(hp_ctx* ctx, const uint8_t id[HP_COMM_ID_BYTES], int rank, int world);
int hp_comm_destroy(hp_ctx* ctx);
/* all-gather of bytes_per_rank bytes from every rank (host buffers; staged through the device, ncclAllGather). */
int hp_comm_allgather(hp_ctx* ctx, const void* send, void* recv, uint64_t bytes_per_rank);
/*
 * Result hand-off: every rank passes the results of its own blocks (local_out, n_local blocks with global indices
 * local_ids[], ascending or not) and receives the results of ALL n_total blocks ordered by global block index, the order
 * OrderedVcfWriter restores for the reference's workers (src/writers/ordered_vcf_writer.rs:158-170).
 * all_var_off[n_total + 1] are the global variant offsets (every rank knows every block's variant count).
 * root < 0: every rank receives (all_out filled everywhere); root >= 0: only that rank copies the gathered message to the
 * host and fills all_out (the other ranks may pass NULL).  all_out->h1, h2 [all_var_off[n_total]], stats, status [n_total];
 * heuristic / counters are not gathered.  Two ncclAllGather calls: the shard sizes, then one message per rank holding its
 * fixed-stride records {index, status, PhaseStats} and its haplotype bytes.
 */
int hp_comm_gather_results(hp_ctx* ctx, uint64_t n_local, const uint64_t* local_ids, const uint64_t* local_var_off,
                           const hp_astar_out* local_out, uint64_t n_total, const uint64_t* all_var_off, int root,
                           hp_astar_out* all_out);

/* ---- post-solve: span counts, block splitting and haplotagging ------------------------
 *      replaces get_solution_span_counts (src/phaser.rs:350-388), the block_split / block_tags loop
 *      (src/phaser.rs:546-569) and haplotag_reads (src/phaser.rs:714-750) ------------------------------------ */
typedef struct hp_post_out {
    uint32_t* span_counts;   /* [n_vars]  block b: N-1 junction counts at var_off[b] .. (last entry of a block unused, 0)   */
    uint64_t* block_tags;    /* [n_vars]  PhaseResult::block_ids: position of the first variant of the variant's sub-block  */
    uint8_t*  read_haplotag; /* [n_reads] 0 / 1, or 2 = unassigned (equal scores: the read gets no entry in the reference)   */
    uint64_t* read_tag;      /* [n_reads] phase block tag of the read (valid when read_haplotag < 2)                         */
} hp_post_out;

/* Host buffers in / out.  var_pos[n_vars] = Variant::position(); h1 / h2 = the haplotypes returned by the A* entry. */
int hp_post_solve_batch(hp_ctx* ctx, const hp_block_batch* batch, const int64_t* var_pos,
                        const uint8_t* h1, const uint8_t* h2, hp_post_out* out);

/* Number of kernel launches this context has issued so far (for bench.py's gpu_launches). */
uint64_t hp_launch_count(const hp_ctx* ctx);
/* Device time (ms, CUDA events on the launching stream) of the dominant kernel in the last _device / _batch call. */
float hp_last_kernel_ms(const hp_ctx* ctx);

/* ---- graph-WFA realignment: replaces WFAGraph::{from_reference_variants_with_hom, edit_distance_with_pruning}
 *      and the traversed-nodes -> allele row glue (src/wfa_graph.rs:119, 350; src/read_parsing.rs:790-851) ---- */
/*
 * Variant table (shared by all jobs of the batch).  Variant k:
 *   position[k], ref_len[k]              Variant::position(), get_ref_len()
 *   allele bytes                         get_truncated_allele0/1 (variants.rs:581-591) at
 *                                        allele_bytes[allele0_off[k] .. allele0_off[k]+allele0_len[k]) etc.
 *   index_allele0[k]                     convert_index(Reference) (variants.rs:649); != 0 means allele0 is an ALT
 *   vtype[k]                             HP_VT_*  (quality table, read_parsing.rs:815-835)
 *   ignored[k]                           is_ignored()
 * Job j aligns read bytes read_bytes[read_off[j] .. read_off[j+1]) against the graph of reference window
 * [ref_start[j], ref_end[j]) (coordinates into `reference`), het variants [het_lo[j], het_hi[j]) and hom
 * variants [hom_lo[j], hom_hi[j]) -- both are index ranges into the one variant table (hets and homs are
 * separate, position-sorted runs of it, exactly the two slices of read_parsing.rs:772-773).
 * Output row j has row_len[j] = het_hi-het_lo cells at alleles[row_off[j] ..] / quals[row_off[j] ..]: the
 * alleles/quals of read_parsing.rs:790-835 restricted to [first_overlap,last_overlap) (everything outside is
 * NoOverlap/0 by construction).
 */
typedef struct hp_variant_table {
    uint32_t        n_variants;
    const int64_t*  position;
    const uint32_t* ref_len;
    const uint64_t* allele0_off;
    const uint32_t* allele0_len;
    const uint64_t* allele1_off;
    const uint32_t* allele1_len;
    const uint8_t*  index_allele0;
    const uint8_t*  vtype;
    const uint8_t*  ignored;
    const uint8_t*  allele_bytes;
    uint64_t        n_allele_bytes;
} hp_variant_table;

typedef struct hp_wfa_batch {
    uint32_t         n_jobs;
    hp_variant_table variants;
    const uint8_t*   reference;      /* chromosome (or concatenated windows) bytes                          */
    uint64_t         n_reference;
    const uint64_t*  ref_start;      /* [n_jobs]                                                            */
    const uint64_t*  ref_end;        /* [n_jobs]  exclusive                                                 */
    const uint32_t*  het_lo;         /* [n_jobs]                                                            */
    const uint32_t*  het_hi;
    const uint32_t*  hom_lo;
    const uint32_t*  hom_hi;
    const uint8_t*   read_bytes;
    const uint64_t*  read_off;       /* [n_jobs+1]                                                          */
    const uint64_t*  row_off;        /* [n_jobs+1] offsets into the output rows                             */
} hp_wfa_batch;

typedef struct hp_wfa_counters {
    uint64_t bases_compared;   /* executions of the compare at wfa_graph.rs:454-456                            */
    uint64_t waves_processed;  /* (offset,set) entries visited at wfa_graph.rs:448                             */
    uint64_t set_ops;          /* set unions / inserts at wfa_graph.rs:491-505, 538-551, 599-613               */
    uint64_t n_nodes;
} hp_wfa_counters;

typedef struct hp_wfa_out {
    int32_t*         status;         /* [n_jobs] HP_WFA_*                                                    */
    uint32_t*        score;          /* [n_jobs] edit distance (max_edit_distance when status==1)            */
    uint8_t*         alleles;        /* [row_off[n_jobs]]                                                    */
    uint8_t*         quals;          /* [row_off[n_jobs]]                                                    */
    uint32_t*        n_nodes;        /* optional [n_jobs]: node count of the job's graph                     */
    uint64_t*        traversed;      /* optional: [n_jobs * trav_words] bitset of traversed node ids         */
    uint32_t         trav_words;     /* 64-bit words per job in `traversed` (0 if traversed == NULL)         */
    hp_wfa_counters* counters;       /* optional [n_jobs]                                                    */
} hp_wfa_out;

/* Host buffers in / out. */
int hp_wfa_align_batch(hp_ctx* ctx, const hp_wfa_batch* batch, hp_wfa_out* out);

/*
 * Pre-built graph variant (the shape of WFAGraph::add_node + edit_distance_with_pruning, wfa_graph.rs:298, 350):
 * node i has bytes seq[seq_off[i] .. seq_off[i+1]) and parents parent_idx[parent_off[i] .. parent_off[i+1]);
 * nodes must be in topological insertion order; the last node is the sink.  One read per call.
 * traversed must hold ceil(n_nodes/64) words.  Returns HP_OK and writes *status = HP_WFA_OK / HP_WFA_MAX_EDIT_DISTANCE.
 */
int hp_wfa_graph_align(hp_ctx* ctx, uint32_t n_nodes, const uint8_t* seq, const uint64_t* seq_off,
                       const uint32_t* parent_idx, const uint64_t* parent_off,
                       const uint8_t* read, uint64_t read_len,
                       uint64_t prune_distance /* UINT64_MAX = off */, uint32_t max_edit_distance,
                       int32_t* status, uint32_t* score, uint64_t* traversed);


/* ---- local realignment: replaces local_realignment (src/read_parsing.rs:121-503),
 *      Variant::match_allele / closest_allele_clip (src/data_types/variants.rs:598-641) and
 *      sequence_alignment::edit_distance (src/sequence_alignment.rs:6-38).  This is the path every read takes when
 *      graph-WFA reports MaxEditDistance or global realignment is switched off for a block (read_parsing.rs:564-600). -- */

/* per-job status written to hp_local_out.status */
#define HP_LOCAL_OK                0
#define HP_LOCAL_UNHANDLED_TYPE    1   /* a reachable variant is SvDuplication / SvInversion / SvBreakend / Unknown:  */
                                       /* the reference panics (read_parsing.rs:320-322, 452-454)                   */
#define HP_LOCAL_BAD_SLICE         2   /* start index after end index in the read: the reference panics on the slice */
#define HP_LOCAL_ALLELE_TOO_LONG   3   /* reserved: no longer reported (comparisons of two sequences beyond 16384     */
                                       /* bases run panel by panel; the reference's grid has no length limit either) */

/* bits of hp_local_out.match_class (the per-variant inputs of ReadStats, read_parsing.rs:457-477) */
#define HP_LOCAL_OVERLAPS  1           /* overlaps_allele                                                           */
#define HP_LOCAL_EXACT     2           /* exact_allele                                                              */

/*
 * Job j = one read mapping (bam::Record) against the variants [var_lo[j], var_hi[j]) of the table.
 *   variants      the FULL alleles (Variant::get_allele0/1: reference prefix + allele + reference postfix, as built by
 *                 src/phaser.rs:236-296) with prefix_len[] / postfix_len[] = get_prefix_len() / get_postfix_len()
 *   read_pos[j]   bam::Record::pos()
 *   segments      the gap-free runs of rust_htslib aligned_pairs() (CIGAR M/=/X): segment s maps reference
 *                 [seg_ref_start[s], +seg_len[s]) onto read [seg_read_start[s], +seg_len[s]); ascending, non-overlapping
 *   read bytes / base qualities at read_bytes[read_off[j] ..], read_quals[read_off[j] ..]
 * Output row j has var_hi-var_lo cells at row_off[j].
 */
typedef struct hp_local_batch {
    uint32_t         n_jobs;
    hp_variant_table variants;        /* index_allele0 is not used by this path and may be NULL                     */
    const uint32_t*  prefix_len;      /* [n_variants]                                                               */
    const uint32_t*  postfix_len;     /* [n_variants]                                                               */
    const uint32_t*  var_lo;          /* [n_jobs]                                                                   */
    const uint32_t*  var_hi;          /* [n_jobs]                                                                   */
    const int64_t*   read_pos;        /* [n_jobs]                                                                   */
    const uint64_t*  seg_off;         /* [n_jobs+1]                                                                 */
    const int64_t*   seg_ref_start;   /* [n_segs]                                                                   */
    const uint32_t*  seg_read_start;  /* [n_segs]                                                                   */
    const uint32_t*  seg_len;         /* [n_segs]                                                                   */
    const uint8_t*   read_bytes;
    const uint8_t*   read_quals;
    const uint64_t*  read_off;        /* [n_jobs+1]                                                                 */
    const uint64_t*  row_off;         /* [n_jobs+1]                                                                 */
} hp_local_batch;

typedef struct hp_local_out {
    uint8_t*  alleles;        /* [row_off[n_jobs]] HP_ALLELE_*                                                      */
    uint8_t*  quals;          /* [row_off[n_jobs]]                                                                  */
    uint8_t*  match_class;    /* optional [row_off[n_jobs]] HP_LOCAL_OVERLAPS | HP_LOCAL_EXACT                      */
    uint32_t* edit_distance;  /* optional [row_off[n_jobs]*2] (d0, d1) of closest_allele_clip where it ran, else 0,0 */
    int32_t*  status;         /* [n_jobs] HP_LOCAL_*                                                                */
} hp_local_out;

/* Host buffers in / out. */
int hp_local_realign_batch(hp_ctx* ctx, const hp_local_batch* batch, hp_local_out* out);

/* sequence_alignment::edit_distance for n_pairs pairs: a = bytes[a_off[i] .. +a_len[i]), b likewise.  Host buffers.
 * Any lengths: pairs whose shorter side exceeds 16384 bases are processed in panels of 16384 pattern rows. */
int hp_edit_distance_batch(hp_ctx* ctx, uint32_t n_pairs, const uint8_t* bytes, uint64_t n_bytes,
                           const uint64_t* a_off, const uint32_t* a_len, const uint64_t* b_off, const uint32_t* b_len,
                           uint32_t* dist);

/* ---- matrix assembly: per-mapping rows -> collapsed ReadSegments -> the A* block batch ------------------------------
 *      replaces ReadSegment::new (src/data_types/read_segments.rs:40-62), ReadSegment::collapse (:71-121), get_num_set
 *      (:151-155) and the min-matched-alleles filter of load_full_read_segments (src/read_parsing.rs:612-629): the glue
 *      between the realignment outputs (hp_wfa_align_batch / hp_local_realign_batch rows) and hp_astar_solve_batch. -- */
#define HP_GROUP_DROPPED   0   /* no allele set after collapsing                                                    */
#define HP_GROUP_PHASABLE  1   /* 0 < num_set < min_matched_alleles: "phasable_segments" only (read_parsing.rs:622-626) */
#define HP_GROUP_KEPT      2   /* becomes a read of the block                                                      */
#define HP_GROUP_ASSERT    3   /* equal alleles with quality 0 on both mappings: the reference asserts (:108)       */
/*
 * Block b owns read groups [group_off[b], group_off[b+1]) (one group per read name, read_parsing.rs:559-562); group g owns
 * rows [group_row_off[g], group_row_off[g+1]) in the order the mappings were pushed.  Row r covers block-relative variant
 * indices [row_start[r], row_start[r] + len) with len = row_cell_off[r+1] - row_cell_off[r]; every other cell of the block
 * is NoOverlap / 0 for that row (the rows of hp_wfa_out / hp_local_out have exactly this shape).
 */
typedef struct hp_rows_batch {
    uint32_t        n_blocks;
    const uint64_t* var_off;          /* [n_blocks+1]                                                               */
    const uint64_t* group_off;        /* [n_blocks+1]                                                               */
    const uint64_t* group_row_off;    /* [n_groups+1]                                                               */
    const uint32_t* row_start;        /* [n_rows]                                                                   */
    const uint64_t* row_cell_off;     /* [n_rows+1]                                                                 */
    const uint8_t*  alleles;          /* [n_row_cells] HP_ALLELE_*                                                  */
    const uint8_t*  quals;            /* [n_row_cells]                                                              */
    uint32_t        min_matched_alleles;  /* --min-matched-alleles (cli.rs), default 2                              */
} hp_rows_batch;

/* Caller-provided buffers: up to n_groups reads and cell_capacity cells (sum over groups of the span from the first to
 * the last row cell always suffices).  read_off .. quals plug straight into hp_block_batch. */
typedef struct hp_assembled {
    uint64_t* read_off;       /* [n_blocks+1]                                                                       */
    uint32_t* read_start;     /* [n_groups]                                                                         */
    uint32_t* read_end;       /* [n_groups]                                                                         */
    uint64_t* cell_off;       /* [n_groups+1]                                                                       */
    uint8_t*  alleles;        /* [cell_capacity]                                                                    */
    uint8_t*  quals;          /* [cell_capacity]                                                                    */
    uint64_t  cell_capacity;
    uint8_t*  group_class;    /* optional [n_groups] HP_GROUP_*                                                     */
    uint32_t* group_num_set;  /* optional [n_groups] ReadSegment::get_num_set of the collapsed segment              */
    uint64_t  n_reads;        /* out                                                                                */
    uint64_t  n_cells;        /* out                                                                                */
} hp_assembled;

/* Host buffers in / out.  Returns HP_ERR_INVALID_INPUT if cell_capacity is too small or a row leaves its block. */
int hp_assemble_blocks(hp_ctx* ctx, const hp_rows_batch* rows, hp_assembled* out);

/* ---- realignment pipeline of a batch of phase blocks: replaces the read loop of load_full_read_segments
 *      (src/read_parsing.rs:545-629).  For every mapping of a block, in BAM order: global realignment (graph-WFA); where it
 *      reports MaxEditDistance the read falls back to local_realignment (:564-575); once a block has seen at least
 *      global_failure_minimum such fallbacks and they make up global_failure_ratio of the reads parsed so far, global
 *      realignment is switched off for the REST of the block (:593-600) -- an order-dependent rule, replayed here exactly;
 *      then ReadSegment::new / collapse per read name / the min-matched-alleles filter (:612-629).  The assembled reads plug
 *      straight into hp_astar_solve_batch.  The GPU runs graph-WFA for all mappings at once, local realignment for the
 *      mappings that need it (two selections: the WFA failures, then everything behind a block's switch-off point). ---- */
#define HP_MAP_GLOBAL          0   /* row of global_realignment                                                        */
#define HP_MAP_LOCAL_FAILED    1   /* graph-WFA reported MaxEditDistance (or HP_WFA_GRAPH_TOO_LARGE): local_realignment */
#define HP_MAP_LOCAL_DISABLED  2   /* global realignment was already switched off for the block: local_realignment     */
#define HP_MAP_SKIPPED         3   /* the mapping overlaps no allele (ReadStats::skipped_reads == 1): no ReadSegment     */

typedef struct hp_realign_batch {
    uint32_t        n_blocks;
    const uint64_t* map_off;          /* [n_blocks+1] mappings of block b (bam records that passed the filter), BAM order */
    const uint32_t* map_group;        /* [n_maps] read-name group of the mapping, block-relative in [0, n_groups[b])      */
    const uint32_t* n_groups;         /* [n_blocks]                                                                       */
    const uint64_t* var_off;          /* [n_blocks+1] het variants of the blocks (= the matrix columns)                   */
    const uint32_t* wfa_het_base;     /* [n_blocks] index of the block's first het variant in wfa.variants                */
    hp_wfa_batch    wfa;              /* job j = mapping j (n_jobs = n_maps)                                              */
    hp_local_batch  local;            /* job j = mapping j; [var_lo, var_hi) = ALL het variants of its block (:568)       */
    uint32_t        global_failure_minimum;   /* --global-failure-count, default 50 (cli.rs:207-212)                     */
    double          global_failure_ratio;     /* --max-global-failure-ratio, default 0.5 (cli.rs:200-205)                */
    uint32_t        min_matched_alleles;      /* --min-matched-alleles, default 2                                        */
} hp_realign_batch;

typedef struct hp_realign_out {
    uint8_t*     map_mode;            /* [n_maps] HP_MAP_*                                                               */
    uint32_t*    map_score;           /* optional [n_maps] wfa_score as the reference records it (:556, :572, :558)       */
    uint32_t*    block_disabled_at;   /* [n_blocks] block-relative index of the mapping that tripped the rule, UINT32_MAX = never */
    uint32_t*    block_failures;      /* optional [n_blocks] num_global_failures at the end of the block                  */
    uint32_t*    block_parsed;        /* optional [n_blocks] total_parsed                                                 */
    hp_assembled assembled;           /* caller-sized: sum n_groups reads; cell_capacity >= sum over groups of the block's n_var */
} hp_realign_out;

/* Host buffers in / out.  A local-realignment job the reference would panic on returns HP_ERR_UNSUPPORTED. */
int hp_realign_block_batch(hp_ctx* ctx, const hp_realign_batch* batch, hp_realign_out* out);

/* ---- CIGAR projection of global realignment on the device: the pre-WFA half of global_realignment
 *      (src/read_parsing.rs:672-742).  For mapping j with aligned segments [seg_off[j], seg_off[j+1]) (the gap-free M/=/X runs
 *      of aligned_pairs(), ascending) against the het / hom calls of its block (positions ascending):
 *        ref_start / ref_end   min_position, max_position + 1                                   (:677-689, :773-774)
 *        het_lo / het_hi       first_overlap, last_overlap: calls with min <= position <= max   (:692-701), table indices
 *        hom_lo / hom_hi       first_hom_overlap (0 -> hom_first when none), last_hom_overlap   (:718-729)
 *        read_start / read_end the read slice aligned against: read[read_start .. read_end)     (:737-741)
 *      het_lo == het_hi marks a mapping that overlaps no het call (the short circuit at :703-712).  These are exactly the
 *      per-job fields of hp_wfa_batch.  A mapping without aligned segments is HP_ERR_INVALID_INPUT (assert at :686). ---- */
typedef struct hp_plan_batch {
    uint32_t        n_maps;
    const uint32_t* map_block;        /* [n_maps] block of the mapping                                                  */
    const uint64_t* seg_off;          /* [n_maps+1]                                                                     */
    const int64_t*  seg_ref_start;    /* [n_segs]                                                                       */
    const uint32_t* seg_read_start;   /* [n_segs]                                                                       */
    const uint32_t* seg_len;          /* [n_segs]                                                                       */
    uint32_t        n_blocks;
    const uint32_t* het_first;        /* [n_blocks+1] het calls of block b = table entries [het_first[b], het_first[b+1]) */
    const int64_t*  het_pos;          /* [het_first[n_blocks]] Variant::position(), ascending inside a block             */
    const uint32_t* hom_first;        /* [n_blocks+1]                                                                   */
    const int64_t*  hom_pos;          /* [hom_first[n_blocks]]                                                          */
} hp_plan_batch;

typedef struct hp_plan_out {
    uint64_t* ref_start;  uint64_t* ref_end;      /* [n_maps]                                                            */
    uint32_t* het_lo;     uint32_t* het_hi;       /* [n_maps] indices into het_pos                                       */
    uint32_t* hom_lo;     uint32_t* hom_hi;       /* [n_maps] indices into hom_pos                                       */
    uint32_t* read_start; uint32_t* read_end;     /* [n_maps]                                                            */
} hp_plan_out;

int hp_wfa_plan_batch(hp_ctx* ctx, const hp_plan_batch* batch, hp_plan_out* out);

/* ---- packed phase-block container + stats writer ----------------------------------------
 * The wire / on-disk form of a block batch, "HPB200" v1 (layout in csrc/hp_pack.cu): what a front end that still owns
 * VCF / BAM decoding (the HiPhase Rust code up to src/phaser.rs:541, or any other reader) writes, and what the batch
 * runner tools/hp_phase_blocks reads.  These functions are host-only (no GPU needed). */
typedef struct hp_packed hp_packed;
/* var_pos may be NULL (then block tags / stats positions fall back to variant indices). */
int  hp_pack_write_blocks(const char* path, const hp_block_batch* batch, const int64_t* var_pos);
/* Reads and validates a whole file; the batch returned by hp_pack_get_blocks points into memory owned by *out. */
int  hp_pack_open(const char* path, hp_packed** out);
int  hp_pack_get_blocks(const hp_packed* packed, hp_block_batch* batch, const int64_t** var_pos);
void hp_pack_close(hp_packed* packed);
const char* hp_pack_last_error(void);
/* One row per block with the solver-side columns of HiPhase's --stats-file (src/writers/phase_stats.rs:207-254);
 * tab separated, comma separated when the path ends in ".csv" (phase_stats.rs:262-271). */
int  hp_write_phase_stats(const char* path, const hp_block_batch* batch, const int64_t* var_pos, const hp_astar_out* out,
                          uint64_t first_block_index);

#ifdef __cplusplus
}
#endif
#endif /* HIPHASE_B200_H */
