// solarosc_batcher.hpp -- rendezvous that turns N single-point callers into one batched launch.
//
// BAT runs its chains on Julia threads, each calling logdensityof(likelihood, params) for ONE parameter point at
// a time (mcmc.jl:96-110, likelihood_main.jl:37-39).  A GPU evaluates thousands of points per launch, so the
// chain threads meet here: every participant deposits its point and blocks; the last one to arrive (or the first
// one whose patience runs out) evaluates the whole batch through the handle and wakes the others.  Participants
// that finish their chain `leave`, after which nobody waits for them.  Host-side C++ only (mutex + condition
// variable); included into solarosc_b200.cu.
#pragma once
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <vector>

struct solarosc_batcher {
    solarosc_batch_fn fn = nullptr;
    void* ctx = nullptr;
    int P = 0, n_slots = 0, want_grad = 0, max_wait_ms = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> active, waiting;     // per slot
    std::vector<int> status;               // per slot: status of the batch that served it
    std::vector<double> params, lnl, grad; // per slot staging
    std::vector<uint32_t> flags;
    std::vector<double> bp, bl, bg;        // packed batch buffers
    std::vector<uint32_t> bf;
    std::vector<int> who;
    unsigned long long generation = 0, n_batches = 0, n_points = 0;
    int n_waiting = 0, n_active = 0;
    bool evaluating = false;

    // called with mu held: evaluate everything deposited so far, publish, wake
    void flush(std::unique_lock<std::mutex>& lk) {
        who.clear();
        for (int s = 0; s < n_slots; ++s)
            if (waiting[s]) who.push_back(s);
        const size_t n = who.size();
        if (n == 0) return;
        bp.resize(n * P); bl.resize(n); bf.resize(n);
        if (want_grad) bg.resize(n * P);
        for (size_t i = 0; i < n; ++i) std::copy(params.begin() + (size_t)who[i] * P, params.begin() + (size_t)(who[i] + 1) * P, bp.begin() + i * P);
        evaluating = true;
        lk.unlock();   // late arrivals may deposit for the next batch meanwhile; they cannot flush while evaluating is set
        const int st = fn(ctx, bp.data(), n, bl.data(), want_grad ? bg.data() : nullptr, bf.data());
        lk.lock();
        evaluating = false;
        for (size_t i = 0; i < n; ++i) {
            const int s = who[i];
            lnl[s] = bl[i]; flags[s] = bf[i]; status[s] = st;
            if (want_grad) std::copy(bg.begin() + i * P, bg.begin() + (i + 1) * P, grad.begin() + (size_t)s * P);
            waiting[s] = 0;
        }
        n_waiting -= (int)n;
        ++generation; ++n_batches; n_points += n;
        cv.notify_all();
    }
};

extern "C" {

solarosc_batcher* solarosc_batcher_create_with(solarosc_batch_fn fn, void* ctx, int n_params, int n_participants, int want_grad,
                                               int max_wait_ms) {
    if (!fn || n_params <= 0 || n_participants <= 0) return nullptr;
    solarosc_batcher* b = new (std::nothrow) solarosc_batcher();
    if (!b) return nullptr;
    b->fn = fn; b->ctx = ctx; b->P = n_params; b->n_slots = n_participants; b->want_grad = want_grad != 0; b->max_wait_ms = max_wait_ms;
    b->active.assign(n_participants, 1); b->waiting.assign(n_participants, 0); b->status.assign(n_participants, 0);
    b->params.assign((size_t)n_participants * n_params, 0.0); b->lnl.assign(n_participants, 0.0);
    b->grad.assign(want_grad ? (size_t)n_participants * n_params : 0, 0.0); b->flags.assign(n_participants, 0u);
    b->n_active = n_participants;
    return b;
}

static int batcher_eval_value(void* ctx, const double* p, size_t n, double* lnl, double* grad, uint32_t* fl) {
    (void)grad;
    return solarosc_eval_llh(static_cast<solarosc_handle*>(ctx), p, n, lnl, fl);
}
static int batcher_eval_grad(void* ctx, const double* p, size_t n, double* lnl, double* grad, uint32_t* fl) {
    return solarosc_eval_llh_grad(static_cast<solarosc_handle*>(ctx), p, n, lnl, grad, fl);
}

solarosc_batcher* solarosc_batcher_create(solarosc_handle* h, int n_participants, int want_grad, int max_wait_ms) {
    if (!h) return nullptr;
    return solarosc_batcher_create_with(want_grad ? batcher_eval_grad : batcher_eval_value, h, solarosc_n_params(h), n_participants,
                                        want_grad, max_wait_ms);
}

int solarosc_batcher_eval(solarosc_batcher* b, int slot, const double* params, double* out_lnl, double* out_grad, uint32_t* out_flag) {
    if (!b || slot < 0 || slot >= b->n_slots || !params || !out_lnl) return SOLAROSC_ERR_ARG;
    std::unique_lock<std::mutex> lk(b->mu);
    if (!b->active[slot] || b->waiting[slot]) return SOLAROSC_ERR_ARG;   // left, or two threads share a slot
    std::copy(params, params + b->P, b->params.begin() + (size_t)slot * b->P);
    b->waiting[slot] = 1;
    ++b->n_waiting;
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(b->max_wait_ms);
    while (b->waiting[slot]) {
        if (b->evaluating) {
            b->cv.wait(lk);                    // a batch is in flight (with or without this slot): its end is signalled
        } else if (b->n_waiting >= b->n_active) {
            b->flush(lk);                      // everybody still in the run is here
        } else if (b->max_wait_ms > 0) {
            if (b->cv.wait_until(lk, deadline) == std::cv_status::timeout && b->waiting[slot] && !b->evaluating)
                b->flush(lk);                  // patience ran out: go with whoever is here
        } else {
            b->cv.wait(lk);
        }
    }
    *out_lnl = b->lnl[slot];
    if (out_flag) *out_flag = b->flags[slot];
    if (out_grad && b->want_grad) std::copy(b->grad.begin() + (size_t)slot * b->P, b->grad.begin() + (size_t)(slot + 1) * b->P, out_grad);
    return b->status[slot];
}

int solarosc_batcher_leave(solarosc_batcher* b, int slot) {
    if (!b || slot < 0 || slot >= b->n_slots) return SOLAROSC_ERR_ARG;
    std::unique_lock<std::mutex> lk(b->mu);
    if (!b->active[slot] || b->waiting[slot]) return SOLAROSC_ERR_ARG;
    b->active[slot] = 0;
    --b->n_active;
    b->cv.notify_all();                        // the others may now be complete without this slot
    return SOLAROSC_OK;
}

int solarosc_batcher_stats(solarosc_batcher* b, unsigned long long* n_batches, unsigned long long* n_points) {
    if (!b) return SOLAROSC_ERR_ARG;
    std::lock_guard<std::mutex> lk(b->mu);
    if (n_batches) *n_batches = b->n_batches;
    if (n_points) *n_points = b->n_points;
    return SOLAROSC_OK;
}

void solarosc_batcher_destroy(solarosc_batcher* b) { delete b; }

}  // extern "C"
