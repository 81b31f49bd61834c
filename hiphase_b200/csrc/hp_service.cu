// hp_service.cu -- the per-block call surface of the reference (astar_solver called from --threads pool workers, one phase block
// each: src/main.rs:385-408, src/phaser.rs:541-543) on top of batched launches: a dispatcher thread packs the blocks of
// concurrent callers into batches and streams them through hp_astar_submit / hp_astar_poll / hp_astar_wait.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hiphase_b200.h"
#include "hp_host.h"

namespace {

struct Request {
    uint32_t n_var, n_reads;
    const uint32_t *read_start, *read_end;
    const uint64_t* cell_off;
    const uint8_t *alleles, *quals, *ignored, *is_snv;
    uint8_t *h1, *h2;
    hp_phase_stats* stats;
    int rc = HP_OK;
    bool done = false;
    std::string err;
};

// One batch in flight: the packed inputs / outputs (pinned, owned by the slot) and the requests it serves.
struct Flight {
    std::vector<Request*> reqs;
    hp::HostPin in, out;
    hp_block_batch batch{};
    hp_astar_out res{};
    hp_astar_job* job = nullptr;
    bool busy = false;
};

}  // namespace

struct hp_service {
    hp_ctx* ctx = nullptr;
    uint32_t max_batch = 4096, linger_us = 200;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Request*> pending;
    bool stop = false;
    std::thread dispatcher;
    std::vector<Flight> flights;
    uint64_t n_batches = 0, n_blocks = 0;
};

namespace {

inline size_t al64(size_t x) { return (x + 63) & ~(size_t)63; }

// Packs the requests into the flight's pinned buffers and submits the batch.  Returns an HP_* code.
int launch(hp_service* s, Flight& f) {
    const size_t nb = f.reqs.size();
    uint64_t nv = 0, nr = 0, nc = 0;
    for (Request* r : f.reqs) { nv += r->n_var; nr += r->n_reads; nc += r->n_reads ? r->cell_off[r->n_reads] - r->cell_off[0] : 0; }
    const size_t in_bytes = al64(8 * (nb + 1)) * 2 + al64(4 * nr) * 2 + al64(8 * (nr + 1)) + al64(nc) * 2 + al64(nv) * 2 + 64;
    const size_t out_bytes = al64(nv) * 2 + al64(sizeof(hp_phase_stats) * nb) + al64(4 * nb) + 64;
    if (!f.in.reserve(in_bytes) || !f.out.reserve(out_bytes)) return HP_ERR_OUT_OF_MEMORY;
    uint8_t* p = (uint8_t*)f.in.ptr;
    auto carve = [&](size_t bytes) { uint8_t* r = p; p += al64(bytes); return r; };
    uint64_t* var_off = (uint64_t*)carve(8 * (nb + 1)); uint64_t* read_off = (uint64_t*)carve(8 * (nb + 1));
    uint32_t* rs = (uint32_t*)carve(4 * nr); uint32_t* re = (uint32_t*)carve(4 * nr);
    uint64_t* co = (uint64_t*)carve(8 * (nr + 1));
    uint8_t* al = carve(nc); uint8_t* ql = carve(nc); uint8_t* ig = carve(nv); uint8_t* sn = carve(nv);
    uint64_t v = 0, r0 = 0, c = 0;
    for (size_t i = 0; i < nb; i++) {
        const Request* q = f.reqs[i];
        var_off[i] = v; read_off[i] = r0;
        if (q->n_reads) {
            memcpy(rs + r0, q->read_start, 4ull * q->n_reads); memcpy(re + r0, q->read_end, 4ull * q->n_reads);
            const uint64_t base = q->cell_off[0], cells = q->cell_off[q->n_reads] - base;
            for (uint32_t k = 0; k < q->n_reads; k++) co[r0 + k] = c + (q->cell_off[k] - base);
            if (cells) { memcpy(al + c, q->alleles + base, cells); memcpy(ql + c, q->quals + base, cells); }
            c += cells;
        }
        if (q->n_var) { memcpy(ig + v, q->ignored, q->n_var); memcpy(sn + v, q->is_snv, q->n_var); }
        v += q->n_var; r0 += q->n_reads;
    }
    var_off[nb] = v; read_off[nb] = r0; co[r0] = c;
    f.batch.n_blocks = (uint32_t)nb; f.batch.var_off = var_off; f.batch.read_off = read_off; f.batch.read_start = rs; f.batch.read_end = re;
    f.batch.cell_off = co; f.batch.alleles = al; f.batch.quals = ql; f.batch.ignored = ig; f.batch.is_snv = sn;
    uint8_t* o = (uint8_t*)f.out.ptr;
    auto ocarve = [&](size_t bytes) { uint8_t* r = o; o += al64(bytes); return r; };
    f.res.h1 = ocarve(nv); f.res.h2 = ocarve(nv);
    f.res.stats = (hp_phase_stats*)ocarve(sizeof(hp_phase_stats) * nb); f.res.status = (int32_t*)ocarve(4 * nb);
    f.res.heuristic = nullptr; f.res.counters = nullptr;
    return hp_astar_submit(s->ctx, &f.batch, &f.res, &f.job);
}

// Results of a finished flight back to the callers (rc per request).
void deliver(hp_service* s, Flight& f, int rc) {
    uint64_t v = 0;
    std::string err = rc != HP_OK ? std::string(hp_last_error(s->ctx)) : std::string();
    for (size_t i = 0; i < f.reqs.size(); i++) {
        Request* q = f.reqs[i];
        if (rc == HP_OK) {
            if (q->n_var) { memcpy(q->h1, f.res.h1 + v, q->n_var); memcpy(q->h2, f.res.h2 + v, q->n_var); }
            if (q->stats) *q->stats = f.res.stats[i];
            if (f.res.status[i] != HP_BLOCK_OK) {      // the reference panics here (astar_phaser.rs:439, 529, 631)
                q->rc = HP_ERR_INVALID_INPUT; q->err = "block rejected with per-block status " + std::to_string(f.res.status[i]);
            }
        } else { q->rc = rc; q->err = err; }
        v += q->n_var;
    }
}

void dispatcher_main(hp_service* s) {
    cudaSetDevice(s->ctx->device);
    std::unique_lock<std::mutex> lk(s->mu);
    for (;;) {
        bool any_busy = false;
        for (Flight& f : s->flights) any_busy |= f.busy;
        if (s->stop && s->pending.empty() && !any_busy) break;
        // a free slot and something to send: linger briefly for company, then pack and submit
        Flight* slot = nullptr;
        for (Flight& f : s->flights) if (!f.busy) { slot = &f; break; }
        if (slot && !s->pending.empty()) {
            if (s->pending.size() < s->max_batch && !any_busy && !s->stop)
                s->cv_work.wait_for(lk, std::chrono::microseconds(s->linger_us), [&] { return s->pending.size() >= s->max_batch || s->stop; });
            slot->reqs.clear();
            while (!s->pending.empty() && slot->reqs.size() < s->max_batch) { slot->reqs.push_back(s->pending.front()); s->pending.pop_front(); }
            lk.unlock();
            const int rc = launch(s, *slot);
            lk.lock();
            if (rc != HP_OK) {
                deliver(s, *slot, rc);
                for (Request* q : slot->reqs) q->done = true;
                s->cv_done.notify_all();
            } else {
                slot->busy = true; s->n_batches++; s->n_blocks += slot->reqs.size();
            }
            continue;
        }
        // poll the flights (oldest first is not required: each has its own lane)
        bool progressed = false;
        for (Flight& f : s->flights) {
            if (!f.busy) continue;
            int done = 0;
            lk.unlock();
            int rc = hp_astar_poll(s->ctx, f.job, &done);
            if (rc == HP_OK && done) rc = hp_astar_wait(s->ctx, f.job);
            else if (rc != HP_OK) { hp_astar_wait(s->ctx, f.job); done = 1; }
            lk.lock();
            if (done) {
                deliver(s, f, rc);
                for (Request* q : f.reqs) q->done = true;
                f.busy = false; f.job = nullptr; progressed = true;
                s->cv_done.notify_all();
            }
        }
        if (progressed) continue;
        if (any_busy) s->cv_work.wait_for(lk, std::chrono::microseconds(50));
        else s->cv_work.wait(lk, [&] { return s->stop || !s->pending.empty(); });
    }
}

}  // namespace

extern "C" {

int hp_service_create(const hp_params* params, int device, uint32_t max_batch_blocks, uint32_t linger_us, hp_service** out) {
    if (!out) return HP_ERR_INVALID_INPUT;
    *out = nullptr;
    hp_ctx* ctx = nullptr;
    int rc = hp_ctx_create(params, device, &ctx);
    if (rc != HP_OK) return rc;
    hp_service* s = new hp_service();
    s->ctx = ctx;
    if (max_batch_blocks) s->max_batch = max_batch_blocks;
    if (linger_us) s->linger_us = linger_us;
    const int lanes = 4;
    hp_ctx_set_lanes(ctx, lanes);
    s->flights.resize(lanes);
    s->dispatcher = std::thread(dispatcher_main, s);
    *out = s;
    return HP_OK;
}

int hp_service_solve_one(hp_service* s, uint32_t n_var, uint32_t n_reads, const uint32_t* read_start, const uint32_t* read_end,
                         const uint64_t* cell_off, const uint8_t* alleles, const uint8_t* quals, const uint8_t* ignored,
                         const uint8_t* is_snv, uint8_t* h1, uint8_t* h2, hp_phase_stats* stats) {
    if (!s || n_var == 0 || !ignored || !is_snv || !h1 || !h2 || (n_reads && (!read_start || !read_end || !cell_off || !alleles || !quals)))
        return HP_ERR_INVALID_INPUT;
    for (uint32_t k = 0; k < n_reads; k++) if (cell_off[k + 1] < cell_off[k]) return HP_ERR_INVALID_INPUT;
    Request r;
    r.n_var = n_var; r.n_reads = n_reads; r.read_start = read_start; r.read_end = read_end; r.cell_off = cell_off;
    r.alleles = alleles; r.quals = quals; r.ignored = ignored; r.is_snv = is_snv; r.h1 = h1; r.h2 = h2; r.stats = stats;
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->stop) return HP_ERR_INVALID_INPUT;
    s->pending.push_back(&r);
    s->cv_work.notify_one();
    s->cv_done.wait(lk, [&] { return r.done; });
    if (r.rc != HP_OK) s->ctx->err = r.err;
    return r.rc;
}

int hp_service_counters(const hp_service* cs, uint64_t* n_batches, uint64_t* n_blocks) {
    hp_service* s = const_cast<hp_service*>(cs);
    if (!s) return HP_ERR_INVALID_INPUT;
    std::lock_guard<std::mutex> lk(s->mu);
    if (n_batches) *n_batches = s->n_batches;
    if (n_blocks) *n_blocks = s->n_blocks;
    return HP_OK;
}

void hp_service_destroy(hp_service* s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop = true;
    }
    s->cv_work.notify_all();
    if (s->dispatcher.joinable()) s->dispatcher.join();
    for (Flight& f : s->flights) { f.in.release(); f.out.release(); }
    hp_ctx_destroy(s->ctx);
    delete s;
}

}  // extern "C"
