// wfa_oracle.cpp -- CPU ORACLE (test infrastructure only, see hp_oracle.h) for the graph-WFA realignment.
//
// Behavioural restatement, same classes of data structure as the reference (hash-map wavefronts keyed by node
// and diagonal, interned dynamic bitsets for the traversed-node sets):
//   WFANode / WFAGraph / add_node               src/wfa_graph.rs:24-68, 298-331
//   from_reference_variants_with_hom            src/wfa_graph.rs:119-284
//   edit_distance_with_pruning                  src/wfa_graph.rs:350-650
//   traversed nodes -> allele / qual row        src/read_parsing.rs:790-851 (quality table :18-22, :815-835)

#include "hp_oracle.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <queue>
#include <random>
#include <thread>
#include <unordered_map>
#include <vector>

struct hpo_graph {
    struct Node { std::vector<uint8_t> seq; std::vector<uint64_t> parents; };
    std::vector<Node> nodes;
    std::vector<std::vector<uint64_t>> edges;
    uint64_t max_edit_distance = 1000;
    // NodeAlleleMap (wfa_graph.rs:19): node -> [(var_index, allele)]
    std::vector<std::pair<uint64_t, std::vector<std::pair<uint64_t, uint8_t>>>> allele_map;
};

namespace {

using Bits = std::vector<uint64_t>;
struct BitsHash {
    size_t operator()(const Bits& b) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (uint64_t w : b) { h ^= w + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};

// wfa_graph.rs:298-331
int64_t add_node(hpo_graph* g, std::vector<uint8_t> seq, std::vector<uint64_t> parents) {
    uint64_t new_index = g->nodes.size();
    if (new_index == 0) { if (!parents.empty()) return -1; }
    else {
        if (parents.empty()) return -1;
        for (uint64_t p : parents) if (new_index <= p) return -1;
    }
    for (uint64_t p : parents) g->edges[p].push_back(new_index);
    std::sort(parents.begin(), parents.end());            // WFANode::new sorts, wfa_graph.rs:36
    g->nodes.push_back(hpo_graph::Node{std::move(seq), std::move(parents)});
    g->edges.emplace_back();
    return (int64_t)new_index;
}

struct VarRef { uint32_t k; int64_t het_index; };  // het_index < 0: hom variant (None)

// wfa_graph.rs:119-284
hpo_graph* graph_from_variants(const hp_wfa_batch* b, uint32_t job, uint64_t max_ed) {
    const hp_variant_table& vt = b->variants;
    const uint8_t* reference = b->reference;
    const uint64_t ref_start = b->ref_start[job], ref_end = b->ref_end[job];
    if (ref_end > b->n_reference || ref_start > ref_end) return nullptr;
    auto* g = new hpo_graph();
    g->max_edit_distance = max_ed;
    auto fail = [&]() { delete g; return (hpo_graph*)nullptr; };
    auto map_insert = [&](uint64_t node, std::vector<std::pair<uint64_t, uint8_t>> v) { g->allele_map.emplace_back(node, std::move(v)); };

    uint64_t previous_end = ref_start;
    int64_t reference_index;
    std::vector<uint64_t> reference_reconnect;
    std::vector<std::pair<uint64_t, uint8_t>> reference_alleles;
    // PriorityQueue<usize, Reverse<usize>>: min reconnect position first (ties: any order, merged below)
    using QE = std::pair<uint64_t, uint64_t>;  // (reconnect, alt_index)
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> reconnect_queue;

    std::vector<VarRef> all;
    for (uint32_t k = b->het_lo[job]; k < b->het_hi[job]; k++) all.push_back(VarRef{k, (int64_t)(k - b->het_lo[job])});
    for (uint32_t k = b->hom_lo[job]; k < b->hom_hi[job]; k++) all.push_back(VarRef{k, -1});
    std::stable_sort(all.begin(), all.end(), [&](const VarRef& a, const VarRef& c) { return vt.position[a.k] < vt.position[c.k]; });

    auto ref_slice = [&](uint64_t s, uint64_t e) { return std::vector<uint8_t>(reference + s, reference + e); };

    for (const VarRef& vr : all) {
        const uint32_t k = vr.k;
        if (vt.ignored[k]) continue;
        if (vt.position[k] < 0) continue;
        const uint64_t variant_pos = (uint64_t)vt.position[k];
        const uint64_t ref_len = vt.ref_len[k];
        if (variant_pos < ref_start) continue;
        if (variant_pos + ref_len > ref_end) continue;

        while (!reconnect_queue.empty() && reconnect_queue.top().first <= variant_pos) {
            auto [alt_reconnect, alt_index] = reconnect_queue.top();
            reconnect_queue.pop();
            if (!(alt_reconnect > previous_end)) return fail();
            reference_index = add_node(g, ref_slice(previous_end, alt_reconnect), reference_reconnect);
            if (reference_index < 0) return fail();
            if (!reference_alleles.empty()) { map_insert(reference_index, reference_alleles); reference_alleles.clear(); }
            previous_end = alt_reconnect;
            reference_reconnect = {(uint64_t)reference_index, alt_index};
            while (!reconnect_queue.empty() && reconnect_queue.top().first == alt_reconnect) {
                reference_reconnect.push_back(reconnect_queue.top().second);
                reconnect_queue.pop();
            }
        }
        if (previous_end < variant_pos || g->nodes.empty()) {
            reference_index = add_node(g, ref_slice(previous_end, variant_pos), reference_reconnect);
            if (reference_index < 0) return fail();
            if (!reference_alleles.empty()) { map_insert(reference_index, reference_alleles); reference_alleles.clear(); }
            reference_reconnect = {(uint64_t)reference_index};
            previous_end = variant_pos;
        } else if (previous_end != variant_pos) return fail();   // assert at wfa_graph.rs:211

        if (vt.index_allele0[k] != 0) {
            std::vector<uint8_t> alt(vt.allele_bytes + vt.allele0_off[k], vt.allele_bytes + vt.allele0_off[k] + vt.allele0_len[k]);
            int64_t alt_index = add_node(g, std::move(alt), reference_reconnect);
            if (alt_index < 0) return fail();
            if (vr.het_index >= 0) map_insert(alt_index, {{(uint64_t)vr.het_index, 0}});
            reconnect_queue.push({variant_pos + ref_len, (uint64_t)alt_index});
        } else if (vr.het_index >= 0) reference_alleles.push_back({(uint64_t)vr.het_index, 0});

        std::vector<uint8_t> alt(vt.allele_bytes + vt.allele1_off[k], vt.allele_bytes + vt.allele1_off[k] + vt.allele1_len[k]);
        int64_t alt_index = add_node(g, std::move(alt), reference_reconnect);
        if (alt_index < 0) return fail();
        if (vr.het_index >= 0) map_insert(alt_index, {{(uint64_t)vr.het_index, 1}});
        reconnect_queue.push({variant_pos + ref_len, (uint64_t)alt_index});
    }
    while (!reconnect_queue.empty()) {
        auto [alt_reconnect, alt_index] = reconnect_queue.top();
        reconnect_queue.pop();
        if (!(alt_reconnect > previous_end)) return fail();
        reference_index = add_node(g, ref_slice(previous_end, alt_reconnect), reference_reconnect);
        if (reference_index < 0) return fail();
        if (!reference_alleles.empty()) { map_insert(reference_index, reference_alleles); reference_alleles.clear(); }
        previous_end = alt_reconnect;
        reference_reconnect = {(uint64_t)reference_index, alt_index};
        while (!reconnect_queue.empty() && reconnect_queue.top().first == alt_reconnect) {
            reference_reconnect.push_back(reconnect_queue.top().second);
            reconnect_queue.pop();
        }
    }
    if (!(previous_end <= ref_end)) return fail();
    if (add_node(g, ref_slice(previous_end, ref_end), reference_reconnect) < 0) return fail();
    if (!reference_alleles.empty()) return fail();   // assert at wfa_graph.rs:281
    return g;
}

using Wave = std::pair<uint64_t, uint64_t>;                         // (offset, set index)
using DiagMap = std::unordered_map<int64_t, std::vector<Wave>>;

// wfa_graph.rs:350-650
int edit_distance(const hpo_graph* g, const uint8_t* other, uint64_t other_len, uint64_t prune_distance,
                  uint64_t* score, std::vector<uint64_t>* traversed, hp_wfa_counters* ctr, uint64_t shuffle_seed) {
    const uint64_t n_nodes = g->nodes.size();
    const size_t words = (n_nodes + 63) / 64;
    std::unordered_map<uint64_t, DiagMap> active, next;
    std::unordered_map<uint64_t, std::unordered_map<int64_t, uint64_t>> max_wavefronts;
    std::unordered_map<Bits, uint64_t, BitsHash> treeset_to_index;
    std::vector<Bits> index_to_treeset;
    std::mt19937_64 rng(shuffle_seed);

    auto intern = [&](Bits&& s) -> uint64_t {
        auto it = treeset_to_index.find(s);
        if (it != treeset_to_index.end()) return it->second;
        index_to_treeset.push_back(s);
        treeset_to_index.emplace(std::move(s), index_to_treeset.size() - 1);
        return index_to_treeset.size() - 1;
    };
    {
        Bits base(words, 0);
        base[0] |= 1ull;
        intern(std::move(base));
    }
    active[0][0].push_back(Wave{0, 0});

    uint64_t ed = 0, farthest = 0, min_progression = 0;
    uint64_t n_cmp = 0, n_waves = 0, n_setops = 0;
    auto finish_counters = [&]() {
        if (ctr) { ctr->bases_compared = n_cmp; ctr->waves_processed = n_waves; ctr->set_ops = n_setops; ctr->n_nodes = n_nodes; }
    };

    for (;;) {
        for (uint64_t node_index = 0; node_index < n_nodes; node_index++) {
            auto ait = active.find(node_index);
            if (ait == active.end()) continue;
            const std::vector<uint8_t>& seq = g->nodes[node_index].seq;
            const uint64_t node_length = seq.size();
            DiagMap wavefront = std::move(ait->second);
            active.erase(ait);
            auto& maxfront = max_wavefronts[node_index];

            std::vector<int64_t> keys;
            keys.reserve(wavefront.size());
            for (auto& kv : wavefront) keys.push_back(kv.first);
            if (shuffle_seed) { std::sort(keys.begin(), keys.end()); std::shuffle(keys.begin(), keys.end(), rng); }

            for (int64_t other_start : keys) {
                std::vector<Wave>& vec_waves = wavefront[other_start];
                uint64_t max_offset = 0;
                for (Wave& w : vec_waves) {
                    n_waves++;
                    uint64_t& offset = w.first;
                    uint64_t other_position = (uint64_t)(other_start + (int64_t)offset);
                    while (offset < node_length && other_position < other_len) {
                        n_cmp++;
                        if (seq[offset] != other[other_position]) break;
                        offset++; other_position++;
                    }
                    max_offset = std::max(max_offset, offset);
                }
                uint64_t& record = maxfront.emplace(other_start, 0).first->second;
                if (max_offset < record || (other_start + (int64_t)max_offset) < (int64_t)min_progression) continue;
                record = max_offset;
                farthest = std::max(farthest, (uint64_t)(other_start + (int64_t)max_offset));

                std::vector<uint64_t> best_sets;
                for (const Wave& w : vec_waves) if (w.first == max_offset) best_sets.push_back(w.second);
                std::sort(best_sets.begin(), best_sets.end());
                best_sets.erase(std::unique(best_sets.begin(), best_sets.end()), best_sets.end());
                uint64_t best_set;
                if (best_sets.size() > 1) {
                    Bits u(words, 0);
                    for (uint64_t si : best_sets) { for (size_t w = 0; w < words; w++) u[w] |= index_to_treeset[si][w]; n_setops++; }
                    best_set = intern(std::move(u));
                } else best_set = best_sets[0];

                if (max_offset == node_length) {
                    if (node_index == n_nodes - 1) {
                        if ((uint64_t)(other_start + (int64_t)max_offset) < other_len)
                            next[node_index][other_start + 1].push_back(Wave{max_offset, best_set});
                    } else {
                        int64_t new_offset = other_start + (int64_t)max_offset;
                        for (uint64_t succ : g->edges[node_index]) {
                            Bits ns = index_to_treeset[best_set];
                            ns[succ / 64] |= 1ull << (succ % 64);
                            n_setops++;
                            uint64_t idx = intern(std::move(ns));
                            active[succ][new_offset].push_back(Wave{0, idx});
                        }
                    }
                } else {
                    DiagMap& node_wf = next[node_index];
                    node_wf[other_start - 1].push_back(Wave{max_offset + 1, best_set});
                    if ((uint64_t)(other_start + (int64_t)max_offset) < other_len) {
                        node_wf[other_start].push_back(Wave{max_offset + 1, best_set});
                        node_wf[other_start + 1].push_back(Wave{max_offset, best_set});
                    }
                }
            }

            if (node_index == n_nodes - 1) {
                std::vector<uint64_t> finals;
                for (auto& kv : wavefront)
                    for (const Wave& w : kv.second)
                        if (w.first == node_length && (uint64_t)(kv.first + (int64_t)w.first) == other_len) finals.push_back(w.second);
                if (!finals.empty()) {
                    std::sort(finals.begin(), finals.end());
                    finals.erase(std::unique(finals.begin(), finals.end()), finals.end());
                    Bits u(words, 0);
                    for (uint64_t si : finals) { for (size_t w = 0; w < words; w++) u[w] |= index_to_treeset[si][w]; }
                    if (finals.size() > 1) n_setops += finals.size();
                    traversed->clear();
                    for (uint64_t i = 0; i < n_nodes; i++) if (u[i / 64] >> (i % 64) & 1) traversed->push_back(i);
                    *score = ed;
                    finish_counters();
                    return HP_WFA_OK;
                }
            }
        }
        ed++;
        active = std::move(next);
        next.clear();
        if (farthest > prune_distance) min_progression = farthest - prune_distance;
        if (ed > g->max_edit_distance) {
            *score = g->max_edit_distance;
            finish_counters();
            return HP_WFA_MAX_EDIT_DISTANCE;
        }
    }
}

// read_parsing.rs:18-22 and :815-835 (global mode doubles the base quality)
int global_qual(uint8_t vtype) {
    switch (vtype) {
        case HP_VT_SNV: return 2 * 80;
        case HP_VT_DELETION: case HP_VT_INSERTION: case HP_VT_INDEL: return 2 * 10;
        case HP_VT_SV_DELETION: case HP_VT_SV_INSERTION: return 2 * 20;
        case HP_VT_TANDEM_REPEAT: return 2 * 40;
        default: return -1;   // panic!("No implementation for matching ...")
    }
}

int align_job(const hp_params* params, const hp_wfa_batch* b, uint32_t j, hp_wfa_out* out) {
    const uint64_t row0 = b->row_off[j], row_len = b->row_off[j + 1] - b->row_off[j];
    const uint64_t n_het = b->het_hi[j] - b->het_lo[j];
    if (row_len != n_het) return -1;
    if (out->alleles) std::memset(out->alleles + row0, HP_ALLELE_NOOVERLAP, row_len);
    if (out->quals) std::memset(out->quals + row0, 0, row_len);
    if (n_het == 0) {                                              // read_parsing.rs:703-712
        out->status[j] = HP_WFA_SKIPPED; out->score[j] = UINT32_MAX;
        return 0;
    }
    hpo_graph* g = graph_from_variants(b, j, params->wfa_max_edit_distance);
    if (!g) return -1;
    uint64_t prune = params->wfa_prune_distance == 0 ? UINT64_MAX : params->wfa_prune_distance;   // cli.rs:352-354
    uint64_t score = 0;
    std::vector<uint64_t> trav;
    hp_wfa_counters ctr{};
    int st = edit_distance(g, b->read_bytes + b->read_off[j], b->read_off[j + 1] - b->read_off[j], prune, &score, &trav, &ctr, 0);
    out->status[j] = st;
    out->score[j] = (uint32_t)score;
    if (out->n_nodes) out->n_nodes[j] = (uint32_t)g->nodes.size();
    if (out->counters) out->counters[j] = ctr;
    if (out->traversed) {
        uint64_t* t = out->traversed + (uint64_t)j * out->trav_words;
        std::memset(t, 0, out->trav_words * 8);
        if (st == HP_WFA_OK) for (uint64_t n : trav) if (n / 64 < out->trav_words) t[n / 64] |= 1ull << (n % 64);
    }
    int rc = 0;
    if (st == HP_WFA_OK) {
        // read_parsing.rs:790-800
        std::vector<uint8_t> alleles(n_het, HP_ALLELE_NOOVERLAP);
        std::unordered_map<uint64_t, const std::vector<std::pair<uint64_t, uint8_t>>*> map;
        for (const auto& kv : g->allele_map) map[kv.first] = &kv.second;
        for (uint64_t n : trav) {
            auto it = map.find(n);
            if (it == map.end()) continue;
            for (const auto& [vi, a] : *it->second) {
                if (alleles[vi] == HP_ALLELE_NOOVERLAP) alleles[vi] = a;
                else if (alleles[vi] != a) alleles[vi] = HP_ALLELE_AMBIGUOUS;
            }
        }
        // read_parsing.rs:803-835
        for (uint64_t i = 0; i < n_het; i++) {
            uint8_t q = 0;
            if (alleles[i] < HP_ALLELE_AMBIGUOUS) {
                int gq = global_qual(b->variants.vtype[b->het_lo[j] + i]);
                if (gq < 0) { rc = -1; gq = 0; }
                q = (uint8_t)gq;
            }
            out->alleles[row0 + i] = alleles[i];
            out->quals[row0 + i] = q;
        }
    }
    delete g;
    return rc;
}

}  // namespace

extern "C" {

hpo_graph* hpo_graph_new(uint64_t max_edit_distance) { auto* g = new hpo_graph(); g->max_edit_distance = max_edit_distance; return g; }
void hpo_graph_free(hpo_graph* g) { delete g; }
int64_t hpo_graph_add_node(hpo_graph* g, const uint8_t* seq, uint64_t len, const uint64_t* parents, uint32_t n_parents) {
    return add_node(g, std::vector<uint8_t>(seq, seq + len), std::vector<uint64_t>(parents, parents + n_parents));
}
uint64_t hpo_graph_num_nodes(const hpo_graph* g) { return g->nodes.size(); }
hpo_graph* hpo_graph_from_job(const hp_wfa_batch* batch, uint32_t job, uint64_t max_edit_distance) {
    return graph_from_variants(batch, job, max_edit_distance);
}
uint64_t hpo_graph_allele_map(const hpo_graph* g, uint64_t* triplets, uint64_t cap) {
    uint64_t n = 0;
    for (const auto& kv : g->allele_map)
        for (const auto& [vi, a] : kv.second) {
            if (n < cap) { triplets[3 * n] = kv.first; triplets[3 * n + 1] = vi; triplets[3 * n + 2] = a; }
            n++;
        }
    return n;
}
void hpo_graph_sizes(const hpo_graph* g, uint64_t* n_nodes, uint64_t* n_seq, uint64_t* n_parents) {
    uint64_t s = 0, p = 0;
    for (const auto& n : g->nodes) { s += n.seq.size(); p += n.parents.size(); }
    *n_nodes = g->nodes.size(); *n_seq = s; *n_parents = p;
}
void hpo_graph_flatten(const hpo_graph* g, uint8_t* seq, uint64_t* seq_off, uint32_t* parent_idx, uint64_t* parent_off) {
    uint64_t s = 0, p = 0;
    for (size_t i = 0; i < g->nodes.size(); i++) {
        seq_off[i] = s; parent_off[i] = p;
        const auto& n = g->nodes[i];
        if (!n.seq.empty()) std::memcpy(seq + s, n.seq.data(), n.seq.size());
        s += n.seq.size();
        for (uint64_t q : n.parents) parent_idx[p++] = (uint32_t)q;
    }
    seq_off[g->nodes.size()] = s; parent_off[g->nodes.size()] = p;
}
int hpo_graph_edit_distance(const hpo_graph* g, const uint8_t* read, uint64_t read_len, uint64_t prune_distance,
                            uint64_t* score, uint64_t* traversed, uint64_t* n_traversed,
                            hp_wfa_counters* counters, uint64_t shuffle_seed) {
    std::vector<uint64_t> trav;
    int st = edit_distance(g, read, read_len, prune_distance, score, &trav, counters, shuffle_seed);
    *n_traversed = trav.size();
    if (traversed) std::copy(trav.begin(), trav.end(), traversed);
    return st;
}
int hpo_wfa_align_batch(const hp_params* params, const hp_wfa_batch* batch, hp_wfa_out* out, int threads) {
    std::atomic<uint32_t> next{0};
    std::atomic<int> failures{0};
    auto worker = [&]() {
        for (;;) {
            uint32_t j = next.fetch_add(1);
            if (j >= batch->n_jobs) break;
            if (align_job(params, batch, j, out) != 0) failures++;
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

}  // extern "C"
