// extern "C" entry points of librbn_b200 (see include/rbn_b200.h for the contract and the reference citations).
#include "common.cuh"
#include <stdarg.h>
#include <mutex>
#include <vector>

namespace rbn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int simt_forward(const RbnDesc*, const WsLayout&, char*, const float*, const float*, const float*, const int*, const int*,
                 float*, float*, float*, const int*, cudaStream_t);
int simt_backward(const RbnDesc*, const WsLayout&, char*, const float*, const float*, const float*, const int*,
                  const int*, const float*, float*, float*, float*, const int*, cudaStream_t);
int export_chart(const RbnDesc*, const WsLayout&, const char*, const int*, int, double*, cudaStream_t);
int simt_viterbi(const RbnDesc*, const WsLayout&, char*, const float*, const float*, const float*, const int*, const int*,
                 float*, int*, int*, cudaStream_t);
int lognormalize(void*, int, long long, long long, int*, cudaStream_t);
int tc_forward(const RbnDesc*, const WsLayout&, char*, const float*, const float*, const float*, const int*, const int*,
               float*, float*, float*, const int*, cudaStream_t);
int tc_backward(const RbnDesc*, const WsLayout&, char*, const float*, const float*, const float*, const int*,
                const int*, const float*, float*, float*, float*, const int*, cudaStream_t);
int debug_gemm(int, int, int, const __half*, const __half*, float*, int, int, int, cudaStream_t);
size_t gauss_workspace_bytes(int, int, int, int);
int gauss_backward_f32(int, int, int, const int*, const double*, const double*, const double*, const double*, const double*, const float*,
                       float*, const double*, float*, double*, double*, double*, double*, double*, double*, cudaStream_t);
int gauss_forward_f32(int, int, int, const float*, const int*, const double*, const double*, const double*, const double*,
                      const double*, const double*, float*, double*, cudaStream_t);
int gauss_backward(int, int, int, const int*, const double*, const double*, const double*, const double*, const double*, const double*,
                   double*, const double*, double*, double*, double*, double*, double*, double*, double*, cudaStream_t);
int gauss_forward(int, int, int, const double*, const int*, const double*, const double*, const double*, const double*,
                  const double*, const double*, double*, double*, cudaStream_t);

// ---- profiling -------------------------------------------------------------------------------------------------------
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static long long g_launches[KC_NUM] = {0};
static double g_ms[KC_NUM] = {0};
struct EvPair { cudaEvent_t a, b; int cls; };
static std::vector<EvPair> g_pending;
static std::vector<EvPair> g_free;

ProfScope::ProfScope(int c, cudaStream_t s) : cls(c), st(s), slot(-1) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_launches[cls]++;
    if (!g_prof_on) return;
    EvPair e;
    if (!g_free.empty()) { e = g_free.back(); g_free.pop_back(); }
    else { cudaEventCreate(&e.a); cudaEventCreate(&e.b); }
    e.cls = cls;
    cudaEventRecord(e.a, st);
    g_pending.push_back(e);
    slot = (int)g_pending.size() - 1;
}
ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_pending[slot].b, st);
}

static bool desc_ok(const RbnDesc* d) {
    if (!d) { set_error("desc is NULL"); return false; }
    if (d->N < 1 || d->V < 1 || d->B < 1 || d->L < 1 || d->n_tvars < 1) {
        set_error("bad descriptor: N=%d V=%d B=%d L=%d n_tvars=%d", d->N, d->V, d->B, d->L, d->n_tvars);
        return false;
    }
    if (d->path < RBN_PATH_AUTO || d->path > RBN_PATH_TCGEN05 || d->flags != 0 || d->chunk_items < 0 || d->cache_bytes < 0) {
        set_error("bad descriptor: path=%d flags=%d chunk_items=%d cache_bytes=%lld", d->path, d->flags, d->chunk_items,
                  (long long)d->cache_bytes);
        return false;
    }
    return true;
}

// The split-sum and child-gradient kernels hold 128 splits of a span per launch (two K-slabs of 64, 128 fix-up threads,
// 128-row stage buffers); wider spans take one launch per window of 128 splits (tc_kernels.cu launch_k1 / launch_k5).
// The cap only bounds the window count (8) and keeps row indices of the chart copies far inside 32 bits.
static constexpr int kTcMaxL = 1025;
static bool tc_shape_ok(const RbnDesc* d) {
    return ((d->N % 64 == 0 && d->N >= 64 && d->N <= 256) || d->N == 512) && d->L >= 2 && d->L <= kTcMaxL;
}

int resolve_path(const RbnDesc* d) {
    if (d->path == RBN_PATH_SIMT) return RBN_PATH_SIMT;
    if (d->path == RBN_PATH_TCGEN05) {
        if (!tc_shape_ok(d)) {
            set_error("tensor-core path needs N in {64,128,192,256,512} and 2 <= L <= %d (got N=%d L=%d)", kTcMaxL, d->N, d->L);
            return RBN_E_UNSUPPORTED;
        }
        return RBN_PATH_TCGEN05;
    }
    return tc_shape_ok(d) ? RBN_PATH_TCGEN05 : RBN_PATH_SIMT;
}

bool make_layout(const RbnDesc* d, WsLayout* o) {
    memset(o, 0, sizeof(*o));
    int path = resolve_path(d);
    if (path < 0) return false;
    size_t N = d->N, B = d->B, L = d->L, C = cells_per_seq(d->L);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off = align_up(off + bytes, 1024); return r; };
    o->chart_val = take(B * C * N * sizeof(float));
    o->chart_exp = take(B * C * sizeof(int));
    if (path != RBN_PATH_TCGEN05) o->alpha = take(B * C * N * sizeof(float));
    o->zroot = take(B * sizeof(float));
    o->rootexp = take(B * sizeof(int));
    if (path == RBN_PATH_TCGEN05) {
        size_t Lp = align_up(L + 32, 16);
        o->Lp = (int)Lp;
        size_t max_items = B * (L - 1);                        // widest diagonal (w = 2)
        // default: whole waves of 128-span tiles on 148 SMs, split-sum buffer Y of at most ~6 GB per slot
        const TcTune& tune = tc_tune();
        size_t wave = 148 * 128;
        size_t chunk = (size_t)(6.0e9 / (2.0 * N * N)) / wave * wave;
        if (chunk < wave) chunk = wave;
        if (d->chunk_items > 0) chunk = (size_t)d->chunk_items;
        if (chunk > max_items) chunk = max_items;
        chunk = align_up(chunk, 128);
        o->chunk = (int)chunk;
        const size_t ns = (size_t)(tune.overlap ? tune.nslot : 1);
        o->nslot = (int)ns;
        // outside chart of the tensor-core path: start-major; the left children of a span are consecutive rows of it, the
        // right children one column of consecutive row blocks (reached through a 3-D tensor map)
        o->alpha = take(B * L * Lp * N * sizeof(float));
        o->lc = take(B * L * Lp * N * sizeof(__half));
        o->rc = take(B * (L + 1) * Lp * N * sizeof(__half));
        o->wt = take(N * N * N * sizeof(__half));
        o->wk = take(N * N * N * sizeof(__half));
        o->wexp = take(N * sizeof(int));
        o->wscale = take(2 * N * sizeof(float));
        o->wdens = take(N * sizeof(float));
        o->wbits = take(2 * N * sizeof(int));
        o->y = take(ns * chunk * N * N * sizeof(__half));
        o->gy = take(ns * chunk * N * N * sizeof(__half));
        o->item_e = take(ns * chunk * sizeof(int));
        o->ghat = take(ns * chunk * N * sizeof(__half));
        o->gabs = take(ns * chunk * N * sizeof(__half));
        o->g32 = take(ns * chunk * N * sizeof(float));
        o->gmax = take(32 * sizeof(int));
        o->gexp = take(ns * chunk * sizeof(int));
        o->acc = take(ns * chunk * N * sizeof(float));
        size_t wide = L > 129 ? B * (L - 129) : 0;             // spans of the widest-but-one window count: diagonal w = 130
        if (wide > chunk) wide = chunk;
        o->ywin = take(wide * N * N * sizeof(__half));
        o->ywin_e = take(wide * sizeof(int));
        // split-sum cache: whole spans only, never more than the batch has, off in the two-stream schedules (their
        // forward and backward passes cut the diagonals differently)
        const size_t per_span = N * N * sizeof(__half) + sizeof(int) + N * sizeof(__half) + N * sizeof(float);
        size_t cap = tune.overlap ? 0 : (size_t)d->cache_bytes / per_span;
        const size_t all_spans = B * (L * (L - 1) / 2);
        if (cap > all_spans) cap = all_spans;
        o->cache_spans = (long long)cap;
        o->ycache = take(cap * N * N * sizeof(__half));
        o->ycache_e = take(cap * sizeof(int));
        o->gabs_all = take(cap * N * sizeof(__half));
        o->g32_all = take(cap * N * sizeof(float));
    }
    o->total = off;
    return true;
}

}  // namespace rbn

using namespace rbn;

extern "C" {

int rbn_abi_version(void) { return RBN_ABI_VERSION; }

const char* rbn_last_error(void) { return g_err; }

int rbn_resolve_path(const RbnDesc* desc) {
    if (!desc_ok(desc)) return RBN_E_BADARG;
    return resolve_path(desc);
}

size_t rbn_workspace_bytes(const RbnDesc* desc) {
    WsLayout ws;
    if (!desc_ok(desc) || !make_layout(desc, &ws)) return 0;
    return ws.total;
}

static bool active_ok(const RbnDesc* d, const int32_t* a) {
    if (!a) return true;
    for (int w = 0; w <= d->L; ++w) {
        if (a[w] < 0 || a[w] > d->B || (w > 0 && a[w] > a[w - 1])) {
            set_error("host_active must be non-increasing with 0 <= active[w] <= B (w = %d: %d)", w, a[w]);
            return false;
        }
    }
    if (a[0] != d->B || a[1] != d->B) {
        set_error("host_active[0] and host_active[1] must equal B (every sequence has length >= 1)");
        return false;
    }
    return true;
}

int rbn_inside_forward_sorted(const RbnDesc* desc, const float* W, const float* T, const float* prior, const int32_t* tokens,
                              const int32_t* lengths, void* workspace, size_t workspace_bytes, float* logZ, float* zroot,
                              float* rootexp, const int32_t* host_active, void* stream) {
    if (!desc_ok(desc)) return RBN_E_BADARG;
    if (!W || !T || !prior || !tokens || !lengths || !workspace || !logZ || !zroot || !rootexp) {
        set_error("rbn_inside_forward: NULL pointer argument");
        return RBN_E_BADARG;
    }
    if (!active_ok(desc, host_active)) return RBN_E_BADARG;
    WsLayout ws;
    if (!make_layout(desc, &ws)) return RBN_E_UNSUPPORTED;
    if (workspace_bytes < ws.total || ((uintptr_t)workspace & 1023)) {
        set_error("workspace too small or misaligned: have %zu need %zu (1024-byte aligned)", workspace_bytes, ws.total);
        return RBN_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (resolve_path(desc) == RBN_PATH_TCGEN05)
        return tc_forward(desc, ws, (char*)workspace, W, T, prior, tokens, lengths, logZ, zroot, rootexp, host_active, st);
    return simt_forward(desc, ws, (char*)workspace, W, T, prior, tokens, lengths, logZ, zroot, rootexp, host_active, st);
}

int rbn_inside_forward(const RbnDesc* desc, const float* W, const float* T, const float* prior, const int32_t* tokens,
                       const int32_t* lengths, void* workspace, size_t workspace_bytes, float* logZ, float* zroot,
                       float* rootexp, void* stream) {
    return rbn_inside_forward_sorted(desc, W, T, prior, tokens, lengths, workspace, workspace_bytes, logZ, zroot, rootexp,
                                     nullptr, stream);
}

int rbn_inside_backward_sorted(const RbnDesc* desc, const float* W, const float* T, const float* prior, const int32_t* tokens,
                               const int32_t* lengths, void* workspace, size_t workspace_bytes, const float* coef, float* gW,
                               float* gT, float* gPrior, const int32_t* host_active, void* stream) {
    if (!desc_ok(desc)) return RBN_E_BADARG;
    if (!W || !T || !prior || !tokens || !lengths || !workspace || !coef || !gW || !gT || !gPrior) {
        set_error("rbn_inside_backward: NULL pointer argument");
        return RBN_E_BADARG;
    }
    if (!active_ok(desc, host_active)) return RBN_E_BADARG;
    WsLayout ws;
    if (!make_layout(desc, &ws)) return RBN_E_UNSUPPORTED;
    if (workspace_bytes < ws.total || ((uintptr_t)workspace & 1023)) {
        set_error("workspace too small or misaligned: have %zu need %zu (1024-byte aligned)", workspace_bytes, ws.total);
        return RBN_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (resolve_path(desc) == RBN_PATH_TCGEN05)
        return tc_backward(desc, ws, (char*)workspace, W, T, prior, tokens, lengths, coef, gW, gT, gPrior, host_active, st);
    return simt_backward(desc, ws, (char*)workspace, W, T, prior, tokens, lengths, coef, gW, gT, gPrior, host_active, st);
}

int rbn_inside_backward(const RbnDesc* desc, const float* W, const float* T, const float* prior, const int32_t* tokens,
                        const int32_t* lengths, void* workspace, size_t workspace_bytes, const float* coef, float* gW,
                        float* gT, float* gPrior, void* stream) {
    return rbn_inside_backward_sorted(desc, W, T, prior, tokens, lengths, workspace, workspace_bytes, coef, gW, gT, gPrior,
                                      nullptr, stream);
}

int rbn_viterbi(const RbnDesc* desc, const float* W, const float* T, const float* prior, const int32_t* tokens,
                const int32_t* lengths, void* workspace, size_t workspace_bytes, float* logP, int32_t* backptr,
                int32_t* root_symbol, void* stream) {
    if (!desc_ok(desc)) return RBN_E_BADARG;
    if (!W || !T || !prior || !tokens || !lengths || !workspace || !logP || !backptr || !root_symbol) {
        set_error("rbn_viterbi: NULL pointer argument");
        return RBN_E_BADARG;
    }
    RbnDesc d = *desc;
    d.path = RBN_PATH_SIMT;                       // the max-semiring kernels exist on the CUDA-core path only
    WsLayout ws;
    if (!make_layout(&d, &ws)) return RBN_E_UNSUPPORTED;
    if (workspace_bytes < ws.total || ((uintptr_t)workspace & 1023)) {
        set_error("workspace too small or misaligned: have %zu need %zu (1024-byte aligned)", workspace_bytes, ws.total);
        return RBN_E_WORKSPACE;
    }
    return simt_viterbi(&d, ws, (char*)workspace, W, T, prior, tokens, lengths, logP, backptr, root_symbol, (cudaStream_t)stream);
}

int rbn_export_chart(const RbnDesc* desc, const int32_t* lengths, const void* workspace, size_t workspace_bytes,
                     int32_t seq, double* out, void* stream) {
    if (!desc_ok(desc)) return RBN_E_BADARG;
    if (!lengths || !workspace || !out || seq < 0 || seq >= desc->B) {
        set_error("rbn_export_chart: bad argument (seq=%d)", seq);
        return RBN_E_BADARG;
    }
    WsLayout ws;
    if (!make_layout(desc, &ws)) return RBN_E_UNSUPPORTED;
    if (workspace_bytes < ws.total) {
        set_error("workspace too small: have %zu need %zu", workspace_bytes, ws.total);
        return RBN_E_WORKSPACE;
    }
    return export_chart(desc, ws, (const char*)workspace, lengths, seq, out, (cudaStream_t)stream);
}

int rbn_lognormalize(float* log_p, int64_t group_size, int64_t groups, int32_t* bad, void* stream) {
    if (!log_p || group_size < 1 || groups < 1) {
        set_error("rbn_lognormalize: bad argument");
        return RBN_E_BADARG;
    }
    return lognormalize(log_p, 0, group_size, groups, bad, (cudaStream_t)stream);
}

int rbn_lognormalize_f64(double* log_p, int64_t group_size, int64_t groups, int32_t* bad, void* stream) {
    if (!log_p || group_size < 1 || groups < 1) {
        set_error("rbn_lognormalize_f64: bad argument");
        return RBN_E_BADARG;
    }
    return lognormalize(log_p, 1, group_size, groups, bad, (cudaStream_t)stream);
}

int rbn_debug_gemm(int32_t M, int32_t N, int32_t K, const void* A, const void* B, float* D, int32_t a_mn, int32_t b_mn,
                   int32_t segmented, void* stream) {
    if (!A || !B || !D || M < 1 || N < 1 || K < 1) {
        set_error("rbn_debug_gemm: bad argument");
        return RBN_E_BADARG;
    }
    return debug_gemm(M, N, K, (const __half*)A, (const __half*)B, D, a_mn, b_mn, segmented, (cudaStream_t)stream);
}

static bool gdesc_ok(const RbnGaussDesc* d) {
    if (!d || d->D < 1 || d->D > 8 || d->B < 1 || d->L < 1 || (d->flags & ~RBN_GAUSS_F32) != 0) {
        set_error("bad Gaussian descriptor");
        return false;
    }
    if ((d->flags & RBN_GAUSS_F32) && d->D != 8) {
        set_error("RBN_GAUSS_F32 (fp32 chart storage) is instantiated for D = 8 (got %d)", d->D);
        return false;
    }
    return true;
}
// the plain entry points take fp64 charts, the _f32 ones fp32 charts: the descriptor has to say the same
static bool gdesc_entry_ok(const RbnGaussDesc* d, int f32, const char* fn) {
    if (!gdesc_ok(d)) return false;
    if (((d->flags & RBN_GAUSS_F32) != 0) != (f32 != 0)) {
        set_error("%s: RbnGaussDesc.flags %s RBN_GAUSS_F32", fn, f32 ? "must carry" : "must not carry");
        return false;
    }
    return true;
}

size_t rbn_gauss_workspace_bytes(const RbnGaussDesc* desc) {
    if (!gdesc_ok(desc)) return 0;
    return 2 * gauss_workspace_bytes(desc->D, desc->B, desc->L, desc->flags & RBN_GAUSS_F32);     // chart + adjoint chart
}

int rbn_gauss_inside_forward(const RbnGaussDesc* desc, const double* y, const int32_t* lengths, const double* cov_term,
                             const double* cov_left, const double* cov_right, const double* prior_mean,
                             const double* prior_cov, const double* log_struct, void* workspace, size_t workspace_bytes,
                             double* logZ, void* stream) {
    if (!gdesc_entry_ok(desc, 0, "rbn_gauss_inside_forward")) return RBN_E_BADARG;
    if (!y || !lengths || !cov_term || !cov_left || !cov_right || !prior_mean || !prior_cov || !log_struct || !workspace || !logZ) {
        set_error("rbn_gauss_inside_forward: NULL pointer argument");
        return RBN_E_BADARG;
    }
    if (workspace_bytes < rbn_gauss_workspace_bytes(desc)) {
        set_error("Gaussian workspace too small");
        return RBN_E_WORKSPACE;
    }
    return gauss_forward(desc->D, desc->B, desc->L, y, lengths, cov_term, cov_left, cov_right, prior_mean, prior_cov,
                         log_struct, (double*)workspace, logZ, (cudaStream_t)stream);
}

int rbn_gauss_inside_backward(const RbnGaussDesc* desc, const double* y, const int32_t* lengths, const double* cov_term,
                              const double* cov_left, const double* cov_right, const double* prior_mean,
                              const double* prior_cov, const double* log_struct, void* workspace, size_t workspace_bytes,
                              const double* coef, double* g_y, double* g_cov_term, double* g_cov_left,
                              double* g_cov_right, double* g_prior_mean, double* g_prior_cov, double* g_log_struct,
                              void* stream) {
    (void)y; (void)cov_term;
    if (!gdesc_entry_ok(desc, 0, "rbn_gauss_inside_backward")) return RBN_E_BADARG;
    if (!lengths || !cov_left || !cov_right || !prior_mean || !prior_cov || !log_struct || !workspace || !coef || !g_y || !g_cov_term ||
        !g_cov_left || !g_cov_right || !g_prior_mean || !g_prior_cov || !g_log_struct) {
        set_error("rbn_gauss_inside_backward: NULL pointer argument");
        return RBN_E_BADARG;
    }
    if (workspace_bytes < rbn_gauss_workspace_bytes(desc)) {
        set_error("Gaussian workspace too small");
        return RBN_E_WORKSPACE;
    }
    double* chart = (double*)workspace;
    double* adj = (double*)((char*)workspace + gauss_workspace_bytes(desc->D, desc->B, desc->L, 0));
    return gauss_backward(desc->D, desc->B, desc->L, lengths, cov_left, cov_right, prior_mean, prior_cov, log_struct, chart, adj, coef,
                          g_y, g_cov_term, g_cov_left, g_cov_right, g_prior_mean, g_prior_cov, g_log_struct,
                          (cudaStream_t)stream);
}

int rbn_gauss_inside_forward_f32(const RbnGaussDesc* desc, const float* y, const int32_t* lengths, const double* cov_term,
                                 const double* cov_left, const double* cov_right, const double* prior_mean,
                                 const double* prior_cov, const double* log_struct, void* workspace, size_t workspace_bytes,
                                 double* logZ, void* stream) {
    if (!gdesc_entry_ok(desc, 1, "rbn_gauss_inside_forward_f32")) return RBN_E_BADARG;
    if (!y || !lengths || !cov_term || !cov_left || !cov_right || !prior_mean || !prior_cov || !log_struct || !workspace || !logZ) {
        set_error("rbn_gauss_inside_forward_f32: NULL pointer argument");
        return RBN_E_BADARG;
    }
    if (workspace_bytes < rbn_gauss_workspace_bytes(desc)) {
        set_error("Gaussian workspace too small");
        return RBN_E_WORKSPACE;
    }
    return gauss_forward_f32(desc->D, desc->B, desc->L, y, lengths, cov_term, cov_left, cov_right, prior_mean, prior_cov,
                             log_struct, (float*)workspace, logZ, (cudaStream_t)stream);
}

int rbn_gauss_inside_backward_f32(const RbnGaussDesc* desc, const float* y, const int32_t* lengths, const double* cov_term,
                                  const double* cov_left, const double* cov_right, const double* prior_mean,
                                  const double* prior_cov, const double* log_struct, void* workspace, size_t workspace_bytes,
                                  const double* coef, float* g_y, double* g_cov_term, double* g_cov_left,
                                  double* g_cov_right, double* g_prior_mean, double* g_prior_cov, double* g_log_struct,
                                  void* stream) {
    (void)y; (void)cov_term;
    if (!gdesc_entry_ok(desc, 1, "rbn_gauss_inside_backward_f32")) return RBN_E_BADARG;
    if (!lengths || !cov_left || !cov_right || !prior_mean || !prior_cov || !log_struct || !workspace || !coef || !g_y || !g_cov_term ||
        !g_cov_left || !g_cov_right || !g_prior_mean || !g_prior_cov || !g_log_struct) {
        set_error("rbn_gauss_inside_backward_f32: NULL pointer argument");
        return RBN_E_BADARG;
    }
    if (workspace_bytes < rbn_gauss_workspace_bytes(desc)) {
        set_error("Gaussian workspace too small");
        return RBN_E_WORKSPACE;
    }
    const float* chart = (const float*)workspace;
    float* adj = (float*)((char*)workspace + gauss_workspace_bytes(desc->D, desc->B, desc->L, 1));
    return gauss_backward_f32(desc->D, desc->B, desc->L, lengths, cov_left, cov_right, prior_mean, prior_cov, log_struct, chart, adj,
                              coef, g_y, g_cov_term, g_cov_left, g_cov_right, g_prior_mean, g_prior_cov, g_log_struct,
                              (cudaStream_t)stream);
}

int rbn_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return RBN_OK;
}

int rbn_profile_read(double* ms, int64_t* launches, int32_t n, int32_t reset) {
    if (!ms || !launches || n < KC_NUM) {
        set_error("rbn_profile_read: need arrays of at least %d entries", (int)KC_NUM);
        return RBN_E_BADARG;
    }
    RBN_CUDA_CHECK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& e : g_pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) g_ms[e.cls] += t;
        g_free.push_back(e);
    }
    g_pending.clear();
    for (int i = 0; i < KC_NUM; ++i) { ms[i] = g_ms[i]; launches[i] = g_launches[i]; }
    if (reset) for (int i = 0; i < KC_NUM; ++i) { g_ms[i] = 0; g_launches[i] = 0; }
    return RBN_OK;
}

}  // extern "C"
