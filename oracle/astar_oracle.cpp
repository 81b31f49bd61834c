// astar_oracle.cpp -- CPU ORACLE (test infrastructure only, see hp_oracle.h) for the A* phaser.
//
// Behavioural restatement of the reference with the same *classes* of data structure (binary heap keyed by the
// three-part priority, interval lookup of active reads, per-node haplotype vector clones, byte-wise rescoring
// of both haplotypes from scratch) so that it is also a fair CPU baseline:
//   ReadSegment            src/data_types/read_segments.rs:19-207
//   AstarNode              src/astar_phaser.rs:13-166
//   PQueueHapTracker       src/astar_phaser.rs:171-231
//   calculate_astar_heuristic  src/astar_phaser.rs:246-292
//   astar_subsolver        src/astar_phaser.rs:311-405
//   astar_solver           src/astar_phaser.rs:426-633
// Reference panics / asserts become OracleError -> non-zero status.

#include "hp_oracle.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

struct OracleError : std::runtime_error {
    int block_status;
    OracleError(int st, const std::string& what) : std::runtime_error(what), block_status(st) {}
};

inline void oassert(bool cond, const char* what, int st = HP_BLOCK_ASSERT) {
    if (!cond) throw OracleError(st, what);
}

// ---------------------------------------------------------------------------------------------------------
// ReadSegment (read_segments.rs:19-62): alleles/quals clipped to region = [first set, last set + 1)
// ---------------------------------------------------------------------------------------------------------
struct ReadSegment {
    std::vector<uint8_t> alleles;
    std::vector<uint8_t> quals;
    size_t start = 0, end = 0;

    // read_segments.rs:40-62
    static ReadSegment make(const uint8_t* a, const uint8_t* q, size_t n) {
        size_t first = n, last = n;
        for (size_t i = 0; i < n; i++) if (a[i] < HP_ALLELE_AMBIGUOUS) { first = i; break; }
        for (size_t i = n; i-- > 0;) if (a[i] < HP_ALLELE_AMBIGUOUS) { last = i + 1; break; }
        ReadSegment rs;
        rs.start = first; rs.end = last;
        if (first < last) {
            rs.alleles.assign(a + first, a + last);
            rs.quals.assign(q + first, q + last);
        }
        return rs;
    }
    // already clipped storage (the packed batch layout)
    static ReadSegment from_clipped(size_t start, size_t end, const uint8_t* a, const uint8_t* q) {
        ReadSegment rs;
        rs.start = start; rs.end = end;
        rs.alleles.assign(a, a + (end - start));
        rs.quals.assign(q, q + (end - start));
        return rs;
    }
    // read_segments.rs:128-143
    uint8_t allele(size_t i) const { return (i >= start && i < end) ? alleles[i - start] : (uint8_t)HP_ALLELE_NOOVERLAP; }
    uint8_t qual(size_t i) const { return (i >= start && i < end) ? quals[i - start] : (uint8_t)0; }
    // read_segments.rs:151-155
    size_t num_set() const {
        size_t c = 0;
        for (uint8_t a : alleles) c += (a < HP_ALLELE_AMBIGUOUS);
        return c;
    }
    // read_segments.rs:177-206
    uint64_t score_partial(const uint8_t* hap, size_t hap_len, size_t offset, uint64_t* cells) const {
        if (hap_len + offset <= start || offset >= end) return 0;
        size_t lo = std::max(start, offset);
        size_t hi = std::min(end, offset + hap_len);
        uint64_t s = 0;
        for (size_t i = lo; i < hi; i++) {
            uint8_t a = allele(i);
            uint8_t h = hap[i - offset];
            if (h < HP_ALLELE_AMBIGUOUS && a != h) s += qual(i);
        }
        if (cells) *cells += (hi - lo);
        return s;
    }
};

// read_segments.rs:71-121; inputs are clipped segments, output is clipped by make()
ReadSegment collapse(const std::vector<ReadSegment>& segs) {
    oassert(!segs.empty(), "collapse: empty", HP_BLOCK_ASSERT);
    if (segs.size() == 1) return segs[0];
    size_t min_start = SIZE_MAX, max_end = 0;
    for (const auto& rs : segs) { min_start = std::min(min_start, rs.start); max_end = std::max(max_end, rs.end); }
    std::vector<uint8_t> alleles(max_end, HP_ALLELE_NOOVERLAP), quals(max_end, 0);
    for (const auto& rs : segs) {
        for (size_t i = min_start; i < max_end; i++) {
            uint8_t rsa = rs.allele(i), rsq = rs.qual(i);
            if (rsa != HP_ALLELE_NOOVERLAP) {
                if (alleles[i] == HP_ALLELE_NOOVERLAP) { alleles[i] = rsa; quals[i] = rsq; }
                else if (alleles[i] == HP_ALLELE_AMBIGUOUS) { /* stays ambiguous, qual 0 */ }
                else if (alleles[i] == rsa) {
                    quals[i] = std::max(quals[i], rsq);
                    oassert(quals[i] > 0, "collapse: equal alleles with zero quality (read_segments.rs:108)");
                } else { alleles[i] = HP_ALLELE_AMBIGUOUS; quals[i] = 0; }
            }
        }
    }
    return ReadSegment::make(alleles.data(), quals.data(), max_end);
}

// ---------------------------------------------------------------------------------------------------------
// Interval lookup: bio::data_structures::interval_tree::IntervalTree<usize, ReadSegment> as used by the
// reference (insert(range, data); find(range) = all entries with start < q.end && end > q.start).
// Static augmented tree over the intervals sorted by start (implicit balanced BST + max_end per subtree).
// ---------------------------------------------------------------------------------------------------------
struct IntervalIndex {
    std::vector<const ReadSegment*> sorted;   // by start
    std::vector<size_t> max_end;              // per implicit subtree [lo,hi) keyed by mid
    void build(const std::vector<ReadSegment>& reads) {
        sorted.clear();
        for (const auto& r : reads) sorted.push_back(&r);
        std::stable_sort(sorted.begin(), sorted.end(), [](const ReadSegment* a, const ReadSegment* b) { return a->start < b->start; });
        max_end.assign(sorted.size(), 0);
        if (!sorted.empty()) fill(0, sorted.size());
    }
    size_t fill(size_t lo, size_t hi) {
        size_t mid = lo + (hi - lo) / 2;
        size_t m = sorted[mid]->end;
        if (lo < mid) m = std::max(m, fill(lo, mid));
        if (mid + 1 < hi) m = std::max(m, fill(mid + 1, hi));
        max_end[mid] = m;
        return m;
    }
    template <class F> void find(size_t qs, size_t qe, F&& f) const {
        if (!sorted.empty()) walk(0, sorted.size(), qs, qe, f);
    }
    template <class F> void walk(size_t lo, size_t hi, size_t qs, size_t qe, F& f) const {
        size_t mid = lo + (hi - lo) / 2;
        if (max_end[mid] <= qs) return;                 // nothing in this subtree ends after qs
        if (lo < mid) walk(lo, mid, qs, qe, f);
        const ReadSegment* r = sorted[mid];
        if (r->start < qe) {
            if (r->end > qs) f(*r);
            if (mid + 1 < hi) walk(mid + 1, hi, qs, qe, f);
        }
    }
};

struct Counters { uint64_t evals = 0, cells = 0, sum_parent_len = 0, pops = 0; };

// ---------------------------------------------------------------------------------------------------------
// AstarNode (astar_phaser.rs:13-166)
// ---------------------------------------------------------------------------------------------------------
struct Priority {              // (Reverse(total), num_hets, Reverse(node_index)), astar_phaser.rs:131-133
    uint64_t cost, hets, index;
    // "a is popped before b"
    bool before(const Priority& o) const {
        if (cost != o.cost) return cost < o.cost;
        if (hets != o.hets) return hets > o.hets;
        return index < o.index;
    }
};

struct AstarNode {
    uint64_t node_index = 0, frozen = 0, fluid = 0, heuristic = 0, num_hets = 0;
    std::vector<uint8_t> h1, h2;
    uint64_t total() const { return frozen + fluid + heuristic; }                        // :126-128
    Priority priority() const { return Priority{total(), num_hets, node_index}; }       // :131-133
    Priority cleared_priority() const { return Priority{0, num_hets, node_index}; }     // :136-138
    size_t allele_count() const { return h1.size(); }
    bool identical() const { return h1 == h2; }                                         // :163-165
};

// astar_phaser.rs:47-57
std::unique_ptr<AstarNode> root_node(uint64_t max_heuristic) {
    auto n = std::make_unique<AstarNode>();
    n->heuristic = max_heuristic;
    return n;
}

// astar_phaser.rs:69-119
std::unique_ptr<AstarNode> extended_node(uint64_t node_index, const AstarNode& parent, uint8_t a1, uint8_t a2,
                                         uint64_t heuristic, const IntervalIndex& reads, size_t hap_offset, Counters* ctr) {
    auto n = std::make_unique<AstarNode>();
    n->h1.reserve(parent.h1.size() + 1); n->h1 = parent.h1; n->h1.push_back(a1);
    n->h2.reserve(parent.h2.size() + 1); n->h2 = parent.h2; n->h2.push_back(a2);
    n->num_hets = parent.num_hets + (a1 == a2 ? 0 : 1);
    uint64_t frozen = parent.frozen, fluid = 0;
    size_t hap_len = n->h1.size() + hap_offset;
    uint64_t cells = 0;
    reads.find(hap_len - 1, hap_len, [&](const ReadSegment& rs) {
        uint64_t c1 = rs.score_partial(n->h1.data(), n->h1.size(), hap_offset, &cells);
        uint64_t c2 = rs.score_partial(n->h2.data(), n->h2.size(), hap_offset, nullptr);
        uint64_t c = std::min(c1, c2);
        if (rs.end <= hap_len) frozen += c; else fluid += c;
    });
    n->node_index = node_index; n->frozen = frozen; n->fluid = fluid; n->heuristic = heuristic;
    if (ctr) { ctr->evals++; ctr->cells += cells; ctr->sum_parent_len += parent.h1.size(); }
    return n;
}

// priority_queue::PriorityQueue<AstarNode, (Reverse<u64>, u64, Reverse<u64>)>: binary heap on the priority.
// The priority is a strict total order (unique node_index) so the pop sequence does not depend on heap internals.
struct NodeQueue {
    struct Entry { Priority pri; std::unique_ptr<AstarNode> node; };
    std::vector<Entry> heap;
    static bool cmp(const Entry& a, const Entry& b) { return b.pri.before(a.pri); }  // std heap = max-heap
    void push(std::unique_ptr<AstarNode> n) {
        Priority p = n->priority();
        heap.push_back(Entry{p, std::move(n)});
        std::push_heap(heap.begin(), heap.end(), cmp);
    }
    const AstarNode& peek() const { return *heap.front().node; }
    Entry pop() {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        Entry e = std::move(heap.back());
        heap.pop_back();
        return e;
    }
    size_t size() const { return heap.size(); }
    bool empty() const { return heap.empty(); }
    // iter_mut() + automatic re-prioritisation (astar_phaser.rs:576-582)
    template <class F> void reprioritize(F&& f) {
        for (auto& e : heap) f(*e.node, e.pri);
        std::make_heap(heap.begin(), heap.end(), cmp);
    }
};

// ---------------------------------------------------------------------------------------------------------
// PQueueHapTracker (astar_phaser.rs:171-231)
// ---------------------------------------------------------------------------------------------------------
struct HapTracker {
    std::vector<size_t> length_counts;
    size_t total_count = 0, threshold = 0;
    explicit HapTracker(size_t max_len) : length_counts(max_len + 1, 0) {}
    void add_hap(size_t v) { length_counts[v]++; if (v >= threshold) total_count++; }
    void remove_hap(size_t v) {
        oassert(length_counts[v] > 0, "tracker remove: count 0");
        length_counts[v]--;
        if (v >= threshold) { oassert(total_count > 0, "tracker remove: total 0"); total_count--; }
    }
    void increase_threshold(size_t t) {
        oassert(t >= threshold, "tracker: threshold decreased");
        for (size_t i = threshold; i < t; i++) total_count -= length_counts[i];
        threshold = t;
    }
    size_t len() const { return total_count; }
};

const uint8_t HAP_ORDER[4][2] = {{0, 1}, {1, 0}, {0, 0}, {1, 1}};   // astar_phaser.rs:367-372, 535-540

// astar_phaser.rs:311-405
std::pair<uint64_t, size_t> astar_subsolver(size_t problem_offset, size_t problem_size, const IntervalIndex& reads,
                                            const std::vector<uint64_t>& H, const std::vector<uint8_t>& bad,
                                            size_t min_queue_size, size_t queue_increment, Counters* ctr) {
    NodeQueue pq;
    uint64_t next_node_index = 1;
    oassert(H[problem_offset] == 0, "subsolver: H[offset] != 0 (astar_phaser.rs:320)");
    pq.push(root_node(H[problem_offset + 1]));
    size_t next_expected = 0;
    uint64_t max_cost_so_far = 0;
    size_t max_visits = min_queue_size + queue_increment * problem_size;
    size_t nodes_visited = 0;

    while (pq.peek().allele_count() < problem_size && nodes_visited < max_visits) {
        NodeQueue::Entry top = pq.pop();
        const AstarNode& top_node = *top.node;
        size_t allele_count = top_node.allele_count();
        nodes_visited++;
        if (ctr) ctr->pops++;
        if (allele_count == next_expected) {
            max_cost_so_far = std::max(max_cost_so_far, top_node.total());
            next_expected++;
        }
        if (bad[problem_offset + allele_count]) {
            auto nn = extended_node(next_node_index, top_node, HP_ALLELE_AMBIGUOUS, HP_ALLELE_AMBIGUOUS,
                                    H[problem_offset + allele_count + 1], reads, problem_offset, ctr);
            next_node_index++;
            oassert(top_node.total() == nn->total(), "subsolver: bad-variant child changed cost (astar_phaser.rs:360)");
            pq.push(std::move(nn));
        } else {
            bool ident = top_node.identical();
            for (const auto& ho : HAP_ORDER) {
                if (!(ho[0] == HP_ALLELE_ALTERNATE && ho[1] == HP_ALLELE_REFERENCE && ident)) {
                    auto nn = extended_node(next_node_index, top_node, ho[0], ho[1],
                                            H[problem_offset + allele_count + 1], reads, problem_offset, ctr);
                    next_node_index++;
                    pq.push(std::move(nn));
                }
            }
        }
    }
    if (pq.peek().allele_count() == problem_size) {
        max_cost_so_far = std::max(max_cost_so_far, pq.peek().total());     // peek, not pop (:397)
        next_expected++;
    }
    return {max_cost_so_far, next_expected - 1};
}

// astar_phaser.rs:246-292
std::vector<uint64_t> calculate_astar_heuristic(size_t num_variants, size_t max_segment_size, const IntervalIndex& reads,
                                                size_t min_queue_size, size_t queue_increment,
                                                const std::vector<uint8_t>& bad, Counters* ctr) {
    oassert(max_segment_size >= 2, "max_segment_size < 2");
    std::vector<uint64_t> H(num_variants + 1, 0);
    size_t max_clip_size = 1;
    for (size_t v = num_variants; v-- > 0;) {
        auto [max_estimate, solve_size] = astar_subsolver(v, max_clip_size, reads, H, bad, min_queue_size / 10, queue_increment, ctr);
        oassert(solve_size >= std::min<size_t>(max_clip_size, 2), "heuristic: solve_size too small (astar_phaser.rs:268)");
        if (bad[v]) H[v] = H[v + 1];
        else {
            oassert(max_estimate >= H[v + 1], "heuristic not monotone (astar_phaser.rs:284)");
            H[v] = max_estimate;
        }
        max_clip_size = std::min(solve_size + 1, max_segment_size);
    }
    return H;
}

struct BlockView {
    size_t n_var;
    std::vector<ReadSegment> reads;
    std::vector<uint8_t> ignored, is_snv;
};

BlockView load_block(const hp_block_batch* b, uint32_t blk) {
    BlockView v;
    v.n_var = (size_t)(b->var_off[blk + 1] - b->var_off[blk]);
    v.ignored.assign(b->ignored + b->var_off[blk], b->ignored + b->var_off[blk + 1]);
    v.is_snv.assign(b->is_snv + b->var_off[blk], b->is_snv + b->var_off[blk + 1]);
    for (uint64_t r = b->read_off[blk]; r < b->read_off[blk + 1]; r++) {
        size_t s = b->read_start[r], e = b->read_end[r];
        oassert(e >= s && e <= v.n_var && b->cell_off[r + 1] - b->cell_off[r] == e - s, "malformed read", HP_BLOCK_ASSERT);
        v.reads.push_back(ReadSegment::from_clipped(s, e, b->alleles + b->cell_off[r], b->quals + b->cell_off[r]));
    }
    return v;
}

struct SolveResult {
    std::vector<uint8_t> h1, h2;
    hp_phase_stats stats{};
    std::vector<uint64_t> H;
    Counters ctr;
};

// astar_phaser.rs:426-633
SolveResult astar_solver(const BlockView& blk, size_t min_queue_size, size_t queue_increment) {
    SolveResult res;
    const size_t num_variants = blk.n_var;
    oassert(num_variants >= 1, "astar_solver: empty block");
    IntervalIndex reads;
    reads.build(blk.reads);

    // :435-442 every ignored variant must be NoOverlap in every read
    for (const auto& rs : blk.reads)
        for (size_t i = 0; i < num_variants; i++)
            if (blk.ignored[i]) oassert(rs.allele(i) == HP_ALLELE_NOOVERLAP, "ignored variant is set in a read", HP_BLOCK_IGNORED_NOT_NOOVERLAP);

    std::vector<uint8_t> bad(blk.ignored.begin(), blk.ignored.end());
    size_t curr_thresh = min_queue_size;
    const size_t max_queue_size = 10 * min_queue_size;                                   // :457
    size_t min_progress = 0;
    NodeQueue pq;
    HapTracker tracker(num_variants);
    size_t next_expected = 0;
    const size_t max_segment_size = 40;                                                  // :466
    std::vector<uint64_t> H = calculate_astar_heuristic(num_variants, max_segment_size, reads, min_queue_size, queue_increment, bad, &res.ctr);

    uint64_t num_pruned = 0;
    const uint64_t estimated_cost = H[0];
    pq.push(root_node(H[0]));
    tracker.add_hap(0);
    uint64_t next_node_index = 1;

    while (pq.peek().allele_count() < num_variants) {
        NodeQueue::Entry top = pq.pop();
        const AstarNode& top_node = *top.node;
        size_t allele_count = top_node.allele_count();
        tracker.remove_hap(allele_count);
        res.ctr.pops++;
        if (allele_count == next_expected) {
            next_expected++;
            if (num_pruned == 0) {
                curr_thresh += queue_increment;
                oassert(curr_thresh == min_queue_size + queue_increment * next_expected, "threshold bookkeeping (astar_phaser.rs:502)");
            }
        }
        if (allele_count < min_progress) {                                              // :507-515
            if (num_pruned == 0) curr_thresh = min_queue_size;
            num_pruned++;
            continue;
        }
        if (bad[allele_count]) {                                                         // :517-531
            auto nn = extended_node(next_node_index, top_node, HP_ALLELE_AMBIGUOUS, HP_ALLELE_AMBIGUOUS,
                                    H[allele_count + 1], reads, 0, &res.ctr);
            next_node_index++;
            oassert(top_node.total() == nn->total(), "bad-variant child changed cost (astar_phaser.rs:529)");
            pq.push(std::move(nn));
            tracker.add_hap(allele_count + 1);
        } else {
            bool ident = top_node.identical();
            for (const auto& ho : HAP_ORDER) {
                if (!(ho[0] == HP_ALLELE_ALTERNATE && ho[1] == HP_ALLELE_REFERENCE && ident)) {
                    auto nn = extended_node(next_node_index, top_node, ho[0], ho[1], H[allele_count + 1], reads, 0, &res.ctr);
                    next_node_index++;
                    pq.push(std::move(nn));
                    tracker.add_hap(allele_count + 1);
                }
            }
        }
        while (tracker.len() > curr_thresh && min_progress < next_expected) {            // :564-585
            min_progress++;
            tracker.increase_threshold(min_progress);
            if (pq.size() > max_queue_size) {
                pq.reprioritize([&](const AstarNode& n, Priority& p) {
                    if (n.allele_count() < min_progress) p = n.cleared_priority();
                });
            }
        }
    }

    NodeQueue::Entry top = pq.pop();
    size_t allele_count = top.node->allele_count();
    tracker.remove_hap(allele_count);
    oassert(allele_count == num_variants, "failed to find solution (astar_phaser.rs:631)");
    res.h1 = top.node->h1;
    res.h2 = top.node->h2;
    hp_phase_stats st{};
    st.pruned_solutions = num_pruned;
    st.estimated_cost = estimated_cost;
    st.actual_cost = top.node->total();
    for (size_t i = 0; i < num_variants; i++) {                                          // :603-615
        if (res.h1[i] != res.h2[i]) { st.phased_variants++; if (blk.is_snv[i]) st.phased_snvs++; }
        else if (res.h1[i] == HP_ALLELE_AMBIGUOUS) st.skipped_variants++;
        else st.homozygous_variants++;
    }
    oassert(st.actual_cost >= st.estimated_cost, "actual < estimated (phase_stats.rs:163)");
    res.stats = st;
    res.H = std::move(H);
    return res;
}

int solve_one_into(const hp_params* params, const hp_block_batch* batch, uint32_t blk, hp_astar_out* out) {
    uint64_t v0 = batch->var_off[blk];
    try {
        BlockView bv = load_block(batch, blk);
        SolveResult r = astar_solver(bv, params->min_queue_size, params->queue_increment);
        std::memcpy(out->h1 + v0, r.h1.data(), r.h1.size());
        std::memcpy(out->h2 + v0, r.h2.data(), r.h2.size());
        out->stats[blk] = r.stats;
        if (out->status) out->status[blk] = HP_BLOCK_OK;
        if (out->heuristic) std::memcpy(out->heuristic + v0 + blk, r.H.data(), r.H.size() * sizeof(uint64_t));
        if (out->counters) {
            out->counters[blk].evals = r.ctr.evals; out->counters[blk].cells = r.ctr.cells;
            out->counters[blk].sum_parent_len = r.ctr.sum_parent_len; out->counters[blk].pops = r.ctr.pops;
        }
        return 0;
    } catch (const OracleError& e) {
        if (out->status) out->status[blk] = e.block_status;
        std::memset(&out->stats[blk], 0, sizeof(hp_phase_stats));
        return 1;
    }
}

}  // namespace

extern "C" {

void hpo_read_segment_region(const uint8_t* alleles, uint64_t n, uint64_t* start, uint64_t* end) {
    std::vector<uint8_t> q(n, 0);
    ReadSegment rs = ReadSegment::make(alleles, q.data(), n);
    *start = rs.start; *end = rs.end;
}

uint64_t hpo_score_partial(uint64_t start, uint64_t end, const uint8_t* alleles, const uint8_t* quals,
                           const uint8_t* hap, uint64_t hap_len, uint64_t offset) {
    ReadSegment rs = ReadSegment::from_clipped(start, end, alleles, quals);
    return rs.score_partial(hap, hap_len, offset, nullptr);
}

int hpo_collapse(uint32_t k, uint64_t n, const uint8_t* alleles, const uint8_t* quals,
                 uint8_t* out_alleles, uint8_t* out_quals, uint64_t* start, uint64_t* end) {
    try {
        std::vector<ReadSegment> segs;
        for (uint32_t i = 0; i < k; i++) segs.push_back(ReadSegment::make(alleles + i * n, quals + i * n, n));
        ReadSegment c = collapse(segs);
        for (uint64_t i = 0; i < n; i++) { out_alleles[i] = c.allele(i); out_quals[i] = c.qual(i); }
        *start = c.start; *end = c.end;
        return 0;
    } catch (const OracleError&) { return 1; }
}

// Matrix assembly (read_parsing.rs:559-562, 612-629): every row becomes a block-length ReadSegment (ReadSegment::new), the
// rows of a read group are collapsed, groups with at least min_matched_alleles set alleles become reads of the block.
int hpo_assemble_blocks(const hp_rows_batch* b, hp_assembled* out) {
    uint64_t n_reads = 0, n_cells = 0;
    out->read_off[0] = 0; out->cell_off[0] = 0;
    for (uint32_t blk = 0; blk < b->n_blocks; blk++) {
        const size_t N = (size_t)(b->var_off[blk + 1] - b->var_off[blk]);
        for (uint64_t g = b->group_off[blk]; g < b->group_off[blk + 1]; g++) {
            std::vector<ReadSegment> segs;
            for (uint64_t r = b->group_row_off[g]; r < b->group_row_off[g + 1]; r++) {
                std::vector<uint8_t> al(N, HP_ALLELE_NOOVERLAP), ql(N, 0);
                const uint64_t c0 = b->row_cell_off[r], len = b->row_cell_off[r + 1] - c0;
                if (b->row_start[r] + len > N) return 2;
                for (uint64_t i = 0; i < len; i++) { al[b->row_start[r] + i] = b->alleles[c0 + i]; ql[b->row_start[r] + i] = b->quals[c0 + i]; }
                segs.push_back(ReadSegment::make(al.data(), ql.data(), N));
            }
            uint8_t cls = HP_GROUP_DROPPED;
            uint32_t nset = 0;
            ReadSegment c;
            if (!segs.empty()) {
                try { c = collapse(segs); nset = (uint32_t)c.num_set(); }
                catch (const OracleError&) { cls = HP_GROUP_ASSERT; }
            }
            if (cls != HP_GROUP_ASSERT) cls = (nset >= b->min_matched_alleles && nset > 0) ? HP_GROUP_KEPT : (nset > 0 ? HP_GROUP_PHASABLE : HP_GROUP_DROPPED);
            if (out->group_class) out->group_class[g] = cls;
            if (out->group_num_set) out->group_num_set[g] = nset;
            if (cls != HP_GROUP_KEPT) continue;
            if (n_cells + (c.end - c.start) > out->cell_capacity) return 3;
            out->read_start[n_reads] = (uint32_t)c.start; out->read_end[n_reads] = (uint32_t)c.end;
            std::copy(c.alleles.begin(), c.alleles.end(), out->alleles + n_cells);
            std::copy(c.quals.begin(), c.quals.end(), out->quals + n_cells);
            n_cells += c.end - c.start;
            n_reads++;
            out->cell_off[n_reads] = n_cells;
        }
        out->read_off[blk + 1] = n_reads;
    }
    out->n_reads = n_reads; out->n_cells = n_cells;
    return 0;
}

int hpo_astar_node_path(const hp_block_batch* b, const uint64_t* H, uint32_t n_steps, const uint8_t* a1, const uint8_t* a2,
                        uint64_t* frozen, uint64_t* total, uint64_t* hets) {
    try {
        BlockView bv = load_block(b, 0);
        IntervalIndex reads; reads.build(bv.reads);
        std::unique_ptr<AstarNode> node = root_node(H[0]);
        for (uint32_t d = 0; d < n_steps; d++) {
            auto nn = extended_node(d + 1, *node, a1[d], a2[d], H[d + 1], reads, 0, nullptr);
            frozen[d] = nn->frozen; total[d] = nn->total(); hets[d] = nn->num_hets;
            node = std::move(nn);
        }
        return 0;
    } catch (const OracleError&) { return 1; }
}

int hpo_tracker_script(uint32_t max_len, uint32_t n_ops, const uint8_t* ops, const uint32_t* vals, uint64_t* lens) {
    try {
        HapTracker t(max_len);
        for (uint32_t i = 0; i < n_ops; i++) {
            if (ops[i] == 0) t.add_hap(vals[i]);
            else if (ops[i] == 1) t.remove_hap(vals[i]);
            else t.increase_threshold(vals[i]);
            lens[i] = t.len();
        }
        return 0;
    } catch (const OracleError&) { return 1; }
}

int hpo_astar_solve_batch(const hp_params* params, const hp_block_batch* batch, hp_astar_out* out, int threads) {
    // worker pool, one block per task, like src/main.rs:332-408
    std::atomic<uint32_t> next{0};
    std::atomic<int> failures{0};
    auto worker = [&]() {
        for (;;) {
            uint32_t b = next.fetch_add(1);
            if (b >= batch->n_blocks) break;
            failures += solve_one_into(params, batch, b, out);
        }
    };
    if (threads <= 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    return failures.load();
}

int hpo_astar_subsolver(const hp_params* params, const hp_block_batch* b, uint64_t problem_offset, uint64_t problem_size,
                        const uint64_t* Hin, uint64_t* max_cost, uint64_t* solved) {
    try {
        BlockView bv = load_block(b, 0);
        IntervalIndex reads; reads.build(bv.reads);
        std::vector<uint64_t> H(Hin, Hin + bv.n_var + 1);
        std::vector<uint8_t> bad(bv.ignored.begin(), bv.ignored.end());
        auto r = astar_subsolver(problem_offset, problem_size, reads, H, bad, params->min_queue_size / 10, params->queue_increment, nullptr);
        *max_cost = r.first; *solved = r.second;
        return 0;
    } catch (const OracleError&) { return 1; }
}

}  // extern "C"
