// Launch counter and optional per-stage CUDA-event timing (used by bench.py for the roofline figures).
// Disabled by default: the hot path records nothing and pays one relaxed atomic load per launch site.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "hvlm_internal.cuh"

namespace hvlm {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_profile_on{0};
static std::mutex g_mu;
struct Rec {
    int stage;
    cudaEvent_t a, b;
};
static std::vector<Rec> g_recs;
static std::vector<cudaEvent_t> g_free;

// Programmatic dependent launch policy.  Measured on B200, 100-frame clip, same box: PDL on every kernel is a loss
// (15.35 vs 15.20 ms) and PDL on the LayerNorm / im2col launches alone is worse still (15.95-16.06 vs 15.67-15.76 ms: their
// many small CTAs get scheduled early and sit in griddepcontrol.wait on slots the GEMM tail needs), but PDL on the GEMMs
// alone wins 2 % (medians 15.03-15.21 vs 15.39-15.51 ms): their prologue (barrier init, TMEM alloc, descriptor prefetch)
// overlaps the tail of the LayerNorm / attention kernel in front of them.  For small batches the launch gaps and
// prologues are a large share of every kernel and PDL everywhere wins: tower forward 1.91 -> 1.58 ms at 1 frame,
// 2.88 -> 2.53 ms at 10, 4.04 -> 3.76 ms at 20.
// Policy: GEMMs always; everything when the caller's hint is set (vit_l14_fwd sets it for <= 32 frames).
// HVLM_PDL=1 / 0 forces all on / off; HVLM_PDL_MASK selects classes for large batches (bit 0 LN/im2col, 1 GEMM, 2 attention).
static thread_local int g_pdl_hint = 0;
bool pdl_enabled(int cls) {
    static const int forced = []() {
        const char* e = getenv("HVLM_PDL");
        return !e ? -1 : (e[0] == '1' ? 1 : 0);
    }();
    // HVLM_PDL_MASK: per-class switch for large batches (bit 0 LayerNorm / im2col, bit 1 GEMM, bit 2 attention)
    static const int mask = []() {
        const char* e = getenv("HVLM_PDL_MASK");
        return e ? atoi(e) : 2;      // default: GEMMs only
    }();
    if (forced >= 0) return forced == 1;
    return g_pdl_hint != 0 || ((mask >> cls) & 1);
}
PdlScope::PdlScope(bool on) : prev_(g_pdl_hint) { g_pdl_hint = on ? 1 : 0; }
PdlScope::~PdlScope() { g_pdl_hint = prev_; }

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static cudaEvent_t get_event() {
    if (!g_free.empty()) {
        cudaEvent_t e = g_free.back();
        g_free.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

StageTimer::StageTimer(int stage, cudaStream_t s) : idx_(-1), s_(s) {
    if (!g_profile_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_mu);
    Rec r{stage, get_event(), get_event()};
    cudaEventRecord(r.a, s);
    g_recs.push_back(r);
    idx_ = static_cast<int>(g_recs.size()) - 1;
}

StageTimer::~StageTimer() {
    if (idx_ < 0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (idx_ < static_cast<int>(g_recs.size())) cudaEventRecord(g_recs[idx_].b, s_);
}

}  // namespace hvlm

extern "C" uint64_t hvlm_launch_count(void) { return hvlm::g_launches.load(std::memory_order_relaxed); }

extern "C" int hvlm_profile_enable(int on) {
    hvlm::g_profile_on.store(on ? 1 : 0, std::memory_order_relaxed);
    return HVLM_OK;
}

extern "C" int hvlm_profile_collect(float* ms_by_stage, int32_t* launches_by_stage) {
    using namespace hvlm;
    if (!ms_by_stage || !launches_by_stage) return HVLM_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < HVLM_STAGE_COUNT; ++i) {
        ms_by_stage[i] = 0.f;
        launches_by_stage[i] = 0;
    }
    int rc = HVLM_OK;
    for (auto& r : g_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) rc = HVLM_ERR_CUDA;
        if (r.stage >= 0 && r.stage < HVLM_STAGE_COUNT) {
            ms_by_stage[r.stage] += ms;
            launches_by_stage[r.stage] += 1;
        }
        g_free.push_back(r.a);
        g_free.push_back(r.b);
    }
    g_recs.clear();
    return rc;
}
