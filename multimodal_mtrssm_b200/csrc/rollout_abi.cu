// C-ABI entry points of librssm_rollout.so (declared in include/rssm_rollout.h).
// Validates sizes, builds kernel arguments and the weight-gradient job lists, launches on the
// caller's stream.  No hidden state; errors are reported through a thread-local message.
#include <stdlib.h>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "wide_common.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return fail("%s: %s", what, cudaGetErrorString(e));
}

bool class_size_ok(int k) { return k == 2 || k == 4 || k == 8 || k == 16; }

int check_mrssm(const RssmMrssmDims* d) {
    if (!d) return fail("dims is NULL");
    if (d->B < 1 || d->T < 1) return fail("B and T must be >= 1 (got B=%d T=%d)", d->B, d->T);
    if (d->D != 32 || d->H != 32 || d->E != 64)
        return fail("unsupported sizes D=%d H=%d E=%d: this build instantiates deter=hidden=32, embed=64", d->D, d->H, d->E);
    if (d->C * d->K != 16 || !class_size_ok(d->K))
        return fail("unsupported distribution_config (class=%d, category=%d): need class*category = 16, class in {2,4,8,16}", d->K,
                    d->C);
    if (d->A < 2 || d->A > 8 || (d->A & 1)) return fail("unsupported action_size %d: need an even size in 2..8", d->A);
    if (d->precision != RSSM_PRECISION_FP32 && d->precision != RSSM_PRECISION_BF16) return fail("bad precision %d", d->precision);
    return 0;
}

int check_mtrssm(const RssmMtrssmDims* d) {
    if (!d) return fail("dims is NULL");
    if (d->B < 1 || d->T < 1) return fail("B and T must be >= 1 (got B=%d T=%d)", d->B, d->T);
    if (d->HD != 32 || d->LD != 32 || d->HH != 32 || d->HR != 32 || d->E != 64)
        return fail("unsupported sizes hd=%d ld=%d head=%d rep=%d E=%d: this build instantiates 32/32/32/32/64", d->HD, d->LD, d->HH,
                    d->HR, d->E);
    if (d->CL * d->KL != 16 || d->CH * d->KH != 16 || !class_size_ok(d->KL) || !class_size_ok(d->KH))
        return fail("unsupported l_dist/h_dist (class,category) = (%d,%d)/(%d,%d): need class*category = 16", d->KL, d->CL, d->KH,
                    d->CH);
    if (d->A < 2 || d->A > 8 || (d->A & 1)) return fail("unsupported action_size %d: need an even size in 2..8", d->A);
    if (!(d->l_tau > 1.f) || !(d->h_tau > 1.f)) return fail("tau must be greater than 1.0 (l_tau=%g h_tau=%g)", d->l_tau, d->h_tau);
    if (d->precision != RSSM_PRECISION_FP32 && d->precision != RSSM_PRECISION_BF16 && d->precision != RSSM_PRECISION_BF16_FUSED)
        return fail("bad precision %d", d->precision);
    if (d->obs_projected != 0 && d->obs_projected != 1) return fail("obs_projected must be 0 or 1 (got %d)", d->obs_projected);
    if (d->obs_projected && d->precision != RSSM_PRECISION_BF16_FUSED)
        return fail("obs_projected (pre-multiplied first-layer partials) is built for RSSM_PRECISION_BF16_FUSED only (got precision %d)", d->precision);
    return 0;
}

// kernel precision policy (fp32-parity / bf16) and saved-record row length of an MMTRSSM precision value
int mt_kernel_precision(int precision) { return precision == RSSM_PRECISION_FP32 ? RSSM_PRECISION_FP32 : RSSM_PRECISION_BF16; }
int mt_saved_ld(int) { return MTRSSM_SAVED_FLOATS; }

#define REQUIRE(ptr)                                                           \
    do {                                                                       \
        if ((ptr) == nullptr) return fail("required pointer %s is NULL", #ptr); \
    } while (0)

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// appends a source segment at staged column `dst` (segments must be added in increasing, gap-free dst order).
// `elem` = element size in bytes of the source (4: fp32, 2: bf16); ptr/ptr0 are element pointers of that type.
void add_seg(rssm::WgradMmaArgs& a, int dst, const void* ptr, int ld, int valid, int elem = 4, int shift = 0,
             const void* ptr0 = nullptr, int ld0 = 0) {
    rssm::WgradSeg& s = a.seg[a.nseg++];
    s.ptr = static_cast<const char*>(ptr), s.ptr0 = static_cast<const char*>(ptr0 ? ptr0 : ptr);
    s.ld_bytes = ld * elem, s.ld0_bytes = (ptr0 ? ld0 : ld) * elem;
    s.valid = valid, s.shift = shift;
    s.c4_begin = dst / 4, s.c4_end = s.c4_begin + (valid + 3) / 4;
    if (elem == 2) {
        s.kind = 2;
    } else {
        const bool vec = (valid % 4 == 0) && (s.ld_bytes % 16 == 0) && (s.ld0_bytes % 16 == 0) && aligned16(s.ptr) && aligned16(s.ptr0);
        s.kind = vec ? 0 : 1;
    }
    a.stride = s.c4_end * 4;
}

// element pointer arithmetic on the opaque records (fp32 or bf16 depending on the precision)
const void* rec_at(const void* base, int elem_off, int elem) { return static_cast<const char*>(base) + (size_t)elem_off * elem; }

void set_out(rssm::WgradMmaArgs& a, int id, float* dW, int ldw, int kvalid, float* db0 = nullptr, float* db1 = nullptr) {
    rssm::WgradOut& o = a.out[id];
    o.dW = dW, o.ldw = ldw, o.kvalid = kvalid, o.db0 = db0, o.db1 = db1;
}


// ---------------------------------------------------------------------------------------------------
// wide MoPoE-MRSSM family (D = H in 64..512): workspace layout + launch sequence (kernels: mrssm_wide_*.cu)
// ---------------------------------------------------------------------------------------------------
bool is_wide(const RssmMrssmDims* d) { return d && d->D != 32; }

int check_mrssm_wide(const RssmMrssmDims* d) {
    if (!d) return fail("dims is NULL");
    if (d->B < 1 || d->T < 1) return fail("B and T must be >= 1 (got B=%d T=%d)", d->B, d->T);
    if (d->D != d->H || d->D % 128 != 0 || d->D < 128 || d->D > 512 || d->E != 64)
        return fail("unsupported sizes D=%d H=%d E=%d: this build instantiates deter=hidden=32 (default family) or "
                    "deter=hidden in {128,256,384,512} (wide family), embed=64", d->D, d->H, d->E);
    if (d->C * d->K != 16 || !class_size_ok(d->K))
        return fail("unsupported distribution_config (class=%d, category=%d): need class*category = 16, class in {2,4,8,16}", d->K,
                    d->C);
    if (d->A < 1 || d->A > 8) return fail("unsupported action_size %d: need 1..8", d->A);
    if (d->precision != RSSM_PRECISION_BF16)
        return fail("the wide family (deter=%d) is built for RSSM_PRECISION_BF16 only (got precision %d)", d->D, d->precision);
    if (d->unimodal) return fail("the unimodal rollout (BaseRSSM.rollout_representation) is built for the default sizes only (deter=%d)", d->D);
    return 0;
}

size_t wide_saved_planes_bytes(const RssmMrssmDims* d) {
    return (size_t)d->T * rssm::wide::NPLANES * ((size_t)((d->B + 127) / 128) * d->D * 128) * 2;
}

struct WideLayout {
    int D, KC, NSL, NBBT, NBBG, ngroups;
    long long plane;  // elements of one record plane = NBBT * D * 128
    size_t pW2, pWhh, pWih, pWhd, pWae, pWve, emb_a, emb_v, h0p, part, rec1, bar, total_fwd;
};

int wide_layout(const RssmMrssmDims* d, bool need_rec1, WideLayout* L) {
    int dev = 0, nsm = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    if (check_cuda(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return 1;
    L->D = d->D, L->KC = d->D / 64, L->NSL = d->D / 32;
    L->NBBT = (d->B + 127) / 128;
    L->NBBG = nsm / L->NSL;
    if (L->NBBG < 1) return fail("device has %d SMs, the wide kernels need at least %d", nsm, L->NSL);
    if (L->NBBG > L->NBBT) L->NBBG = L->NBBT;
    L->ngroups = (L->NBBT + L->NBBG - 1) / L->NBBG;
    L->plane = (long long)L->NBBT * d->D * 128;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = rssm::wide::align_up(o + bytes, 256); return at; };
    const size_t DD = (size_t)d->D * d->D * 2;
    L->pW2 = take(DD), L->pWhh = take(3 * DD), L->pWih = take(3 * DD), L->pWhd = take(3 * DD);
    L->pWae = take((size_t)d->D * 64 * 2), L->pWve = take((size_t)d->D * 64 * 2);
    const size_t emb = (size_t)d->T * L->NBBT * 64 * 128 * 2;
    L->emb_a = take(emb), L->emb_v = take(emb);
    L->h0p = take((size_t)L->plane * 2);
    L->part = take((size_t)L->NBBT * 128 * L->NSL * 48 * 4);
    L->rec1 = take(need_rec1 ? (size_t)rssm::wide::NPLANES * L->plane * 2 : 0);
    L->bar = take(256 * (size_t)(L->NBBT + 1) + 4096);  // one barrier counter line per batch block, status word, 4 KB of phase timestamps (debug)
    L->total_fwd = o;
    return 0;
}

struct WideBwdLayout {
    size_t pW1x, pWhdT, pWgT, pWihTn, pWhhTn, pW2T, pWaeT, pWveT, ones, drec, stat, h0p, emb_a, emb_v, tiles, bar, total;
    long long dt_stride, dlg, xin;  // elements: gradient-plane step, narrow planes inside a step
};
constexpr int MAX_WIDE_TILES = rssm::MAX_WIDE_WGRAD_TILES;

void wide_bwd_layout(const RssmMrssmDims* d, const WideLayout& L, WideBwdLayout* W) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = rssm::wide::align_up(o + bytes, 256); return at; };
    const size_t D = d->D, DD = D * D * 2;
    W->pW1x = take(32 * D * 2), W->pWhdT = take(3 * DD), W->pWgT = take(4 * DD), W->pWihTn = take(DD), W->pWhhTn = take(DD);
    W->pW2T = take(DD), W->pWaeT = take(64 * D * 2), W->pWveT = take(64 * D * 2), W->ones = take(16 * 128 * 8 * 2);
    W->dlg = (long long)rssm::wide::NDPLANES * L.plane;
    W->xin = W->dlg + (long long)L.NBBT * 48 * 128;
    W->dt_stride = W->xin + (long long)L.NBBT * 32 * 128;
    W->drec = take((size_t)d->T * W->dt_stride * 2);
    W->stat = take((size_t)d->T * L.NBBT * 32 * 128 * 16);  // per-row statistics of the pre-pass
    W->h0p = take((size_t)L.plane * 2);
    const size_t emb = (size_t)d->T * L.NBBT * 64 * 128 * 2;
    W->emb_a = take(emb), W->emb_v = take(emb);
    W->tiles = take(sizeof(rssm::WideWgradTile) * MAX_WIDE_TILES);
    W->bar = take(256 * (size_t)(L.NBBT + 1) + 4096);  // counter line per batch block, status, phase timestamps (debug)
    W->total = o;
}

size_t wide_bwd_workspace(const RssmMrssmDims* d, const WideLayout& L) {
    WideBwdLayout W;
    wide_bwd_layout(d, L, &W);
    return W.total;
}


// CTAs per cluster along the slice index (the activation chunks are multicast inside a cluster).  Default 1: measured on B200
// at cfg3, clusters of 2 / 4 are correct but SLOWER (6.7 ms against 6.2-6.3 ms): the operand ring is 4-5 stages deep and every slot
// release then waits for the slowest of the cluster's CTAs, which costs more than the L2 reads it saves.  RSSM_WIDE_CLUSTER=2|4
// selects them for experiments.
int wide_cluster_size(int NSL) {
    int cs = 1;
    if (const char* e = getenv("RSSM_WIDE_CLUSTER")) cs = atoi(e);
    if (cs != 1 && cs != 2 && cs != 4) cs = 1;
    while (cs > 1 && NSL % cs != 0) cs >>= 1;
    return cs;
}

int wide_pack_weights(const RssmMrssmDims* d, const RssmMrssmWeights* w, char* ws, const WideLayout& L, bool imagine, cudaStream_t s) {
    const int D = d->D;
    rssm::WidePackJobs J{};
    J.NSL = L.NSL;
    auto job = [&](const float* a, int lda, const float* b, int ldb, const float* c, int ldc, int nparts, int coloff, int K, size_t off) {
        rssm::WidePackJob& j = J.job[J.njobs++];
        j.src[0] = a, j.src[1] = b, j.src[2] = c, j.ld[0] = lda, j.ld[1] = ldb, j.ld[2] = ldc;
        j.nparts = nparts, j.coloff = coloff, j.K = K, j.dst = reinterpret_cast<__nv_bfloat16*>(ws + off);
    };
    job(w->asp_w2, D, nullptr, 0, nullptr, 0, 1, 0, D, L.pW2);
    job(w->w_hh, D, w->w_hh + (size_t)D * D, D, w->w_hh + 2 * (size_t)D * D, D, 3, 0, D, L.pWhh);
    job(w->w_ih, D, w->w_ih + (size_t)D * D, D, w->w_ih + 2 * (size_t)D * D, D, 3, 0, D, L.pWih);
    if (imagine) {
        job(w->pr_w1, D, w->pr_w1, D, w->pr_w1, D, 3, 0, D, L.pWhd);
    } else {
        job(w->pr_w1, D, w->au_w1, D + 64, w->vi_w1, D + 64, 3, 0, D, L.pWhd);
        job(w->au_w1, D + 64, nullptr, 0, nullptr, 0, 1, D, 64, L.pWae);
        job(w->vi_w1, D + 64, nullptr, 0, nullptr, 0, 1, D, 64, L.pWve);
    }
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_wide_pack_weights(J, s), "wide weight packing launch");
}

int wide_mrssm_fwd(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in, const RssmMrssmOutputs* out,
                   cudaStream_t s, bool imagine) {
    if (check_mrssm_wide(d)) return 1;
    const bool save = !imagine && out->saved != nullptr;
    WideLayout L;
    if (wide_layout(d, true, &L)) return 1;  // same layout as rssm_mrssm_workspace_bytes reports (the one-step record is unused when saving)
    if (out->workspace == nullptr || out->workspace_bytes < L.total_fwd)
        return fail("wide family: workspace of %zu bytes required (rssm_mrssm_workspace_bytes), got %zu", L.total_fwd,
                    out->workspace ? out->workspace_bytes : (size_t)0);
    if ((reinterpret_cast<uintptr_t>(out->workspace) & 255) != 0) return fail("workspace must be 256-byte aligned");
    char* ws = static_cast<char*>(out->workspace);
    const int D = d->D, T = d->T, A = d->A, F = D + 16, C = d->C;
    // pad rows of the last batch block take part in the contractions: keep them finite
    if (d->B % 128 != 0) {
        if (check_cuda(cudaMemsetAsync(ws + L.emb_a, 0, L.bar - L.emb_a, s), "workspace memset")) return 1;
        if (save && check_cuda(cudaMemsetAsync(out->saved, 0, (size_t)T * rssm::wide::NPLANES * L.plane * 2, s), "record memset")) return 1;
    }
    if (check_cuda(cudaMemsetAsync(ws + L.bar, 0, 256 * (size_t)(L.NBBT + 1) + 4096, s), "barrier memset")) return 1;
    if (wide_pack_weights(d, w, ws, L, imagine, s)) return 1;
    if (!imagine) {
        g_launches.fetch_add(2);
        if (check_cuda(rssm::launch_wide_pack_rows(in->embed_a, d->B, T, 64, 64, 0, reinterpret_cast<__nv_bfloat16*>(ws + L.emb_a), L.NBBT, s),
                       "embedding packing launch"))
            return 1;
        if (check_cuda(rssm::launch_wide_pack_rows(in->embed_v, d->B, T, 64, 64, 0, reinterpret_cast<__nv_bfloat16*>(ws + L.emb_v), L.NBBT, s),
                       "embedding packing launch"))
            return 1;
    }
    const char* swap_env = getenv("RSSM_WIDE_DESC_SWAP");
    for (int g = 0; g < L.ngroups; ++g) {
        const int bb0 = g * L.NBBG, nbb = (L.NBBT - bb0 < L.NBBG) ? L.NBBT - bb0 : L.NBBG;
        const long long r0 = (long long)bb0 * 128;
        rssm::MrssmWideFwdArgs a{};
        a.B = (int)((d->B - r0 < (long long)nbb * 128) ? d->B - r0 : (long long)nbb * 128);
        a.T = T, a.A = A, a.K = d->K, a.D = D, a.NBB = nbb, a.NSL = L.NSL, a.imagine = imagine ? 1 : 0;
        a.desc_swap = (swap_env && swap_env[0] == '1') ? 1 : 0;
        a.plane_stride = L.plane, a.t_stride = save ? (long long)rssm::wide::NPLANES * L.plane : 0;
        a.emb_t_stride = (long long)L.NBBT * 64 * 128;
        a.w = *w;
        auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
        a.pW2 = bf(L.pW2), a.pWhh = bf(L.pWhh), a.pWih = bf(L.pWih), a.pWhd = bf(L.pWhd), a.pWae = bf(L.pWae), a.pWve = bf(L.pWve);
        a.actions = in->actions + r0 * T * A, a.h0 = in->h0 + r0 * D, a.z0 = in->z0 + r0 * 16;
        a.u_post = in->u_post ? in->u_post + r0 * T * C : nullptr, a.u_prior = in->u_prior ? in->u_prior + r0 * T * C : nullptr;
        a.emb_a = bf(L.emb_a) + (long long)bb0 * 64 * 128, a.emb_v = bf(L.emb_v) + (long long)bb0 * 64 * 128;
        a.feature = out->feature + r0 * T * F, a.prior_probs = out->prior_probs + r0 * T * 16;
        a.post_probs = out->post_probs ? out->post_probs + r0 * T * 16 : nullptr;
        a.prior_stoch = out->prior_stoch ? out->prior_stoch + r0 * T * 16 : nullptr;
        a.kl = out->kl ? out->kl + r0 * T : nullptr;
        a.rec = (save ? static_cast<__nv_bfloat16*>(out->saved) : bf(L.rec1)) + (long long)bb0 * D * 128;
        a.h0p = bf(L.h0p) + (long long)bb0 * D * 128;
        a.part = reinterpret_cast<float*>(ws + L.part) + r0 * L.NSL * 48;
        a.logits = save ? reinterpret_cast<float*>(static_cast<char*>(out->saved) + wide_saved_planes_bytes(d)) + r0 * T * 32 : nullptr;
        a.bar = reinterpret_cast<unsigned*>(ws + L.bar + 256 * (size_t)bb0);
        a.status = reinterpret_cast<int*>(ws + L.bar + 256 * (size_t)L.NBBT);
        a.timing = (g == 0 && getenv("RSSM_WIDE_TIMING")) ? reinterpret_cast<unsigned long long*>(ws + L.bar + 256 * (size_t)(L.NBBT + 1)) : nullptr;
        a.exp = getenv("RSSM_WIDE_EXP") ? atoi(getenv("RSSM_WIDE_EXP")) : 0;
        a.cs = wide_cluster_size(L.NSL);
        g_launches.fetch_add(1);
        if (check_cuda(rssm::launch_mrssm_wide_fwd(a, s), imagine ? "wide mrssm imagine launch" : "wide mrssm forward launch")) return 1;
    }
    return 0;
}

int wide_mrssm_bwd(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in, const RssmMrssmOutputs* fo,
                   const RssmMrssmUpstream* up, const RssmMrssmInputGrads* gin, const RssmMrssmWeightGrads* gw, cudaStream_t s) {
    if (check_mrssm_wide(d)) return 1;
    WideLayout L;
    if (wide_layout(d, false, &L)) return 1;
    WideBwdLayout W;
    wide_bwd_layout(d, L, &W);
    if (gin->workspace == nullptr || gin->workspace_bytes < W.total)
        return fail("wide family: backward workspace of %zu bytes required (rssm_mrssm_workspace_bytes(dims, 1)), got %zu", W.total,
                    gin->workspace ? gin->workspace_bytes : (size_t)0);
    if ((reinterpret_cast<uintptr_t>(gin->workspace) & 255) != 0) return fail("workspace must be 256-byte aligned");
    char* ws = static_cast<char*>(gin->workspace);
    auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
    const int D = d->D, T = d->T, A = d->A, F = D + 16;
    const long long plane = L.plane, tstride = (long long)rssm::wide::NPLANES * plane;
    if (d->B % 128 != 0 && check_cuda(cudaMemsetAsync(ws + W.drec, 0, W.tiles - W.drec, s), "workspace memset")) return 1;
    if (check_cuda(cudaMemsetAsync(ws + W.bar, 0, 256 * (size_t)(L.NBBT + 1) + 4096, s), "barrier memset")) return 1;
    const __nv_bfloat16* rec = static_cast<const __nv_bfloat16*>(fo->saved);
    const float* logits = reinterpret_cast<const float*>(static_cast<const char*>(fo->saved) + wide_saved_planes_bytes(d));

    rssm::MrssmWideBwdArgs a{};
    a.T = T, a.A = A, a.K = d->K, a.D = D, a.NSL = L.NSL, a.kl_wq = up->kl_wq, a.kl_wp = up->kl_wp;
    a.plane_stride = plane, a.t_stride = tstride, a.dt_stride = W.dt_stride;
    a.w = *w;
    a.pW1x = bf(W.pW1x), a.pWhdT = bf(W.pWhdT), a.pWgT = bf(W.pWgT), a.pWihTn = bf(W.pWihTn), a.pWhhTn = bf(W.pWhhTn), a.pW2T = bf(W.pW2T);
    g_launches.fetch_add(2);
    if (check_cuda(rssm::launch_wide_pack_bwd_weights(a, s), "wide backward weight packing launch")) return 1;
    if (check_cuda(rssm::launch_wide_pack_dembed(w->au_w1, w->vi_w1, D, bf(W.pWaeT), bf(W.pWveT), bf(W.ones), s), "wide embed-weight packing launch"))
        return 1;
    {
        rssm::WideRowstatArgs r{};
        r.B = d->B, r.T = T, r.A = A, r.K = d->K, r.D = D, r.NBBT = L.NBBT, r.kl_wq = up->kl_wq, r.kl_wp = up->kl_wp, r.dt_stride = W.dt_stride;
        r.logits = logits, r.feature = fo->feature, r.prior_probs = fo->prior_probs, r.post_probs = fo->post_probs, r.z0 = in->z0;
        r.actions = in->actions, r.d_feature = up->d_feature, r.d_prior_probs = up->d_prior_probs, r.d_post_probs = up->d_post_probs;
        r.d_prior_stoch = up->d_prior_stoch, r.d_kl = up->d_kl;
        r.stat = reinterpret_cast<float*>(ws + W.stat), r.xin = bf(W.drec) + W.xin;
        g_launches.fetch_add(1);
        if (check_cuda(rssm::launch_wide_bwd_rowstat(r, s), "wide backward pre-pass launch")) return 1;
    }
    for (int g = 0; g < L.ngroups; ++g) {
        const int bb0 = g * L.NBBG, nbb = (L.NBBT - bb0 < L.NBBG) ? L.NBBT - bb0 : L.NBBG;
        const long long r0 = (long long)bb0 * 128, boff = (long long)bb0 * D * 128;
        a.B = (int)((d->B - r0 < (long long)nbb * 128) ? d->B - r0 : (long long)nbb * 128);
        a.NBB = nbb;
        a.rec = rec + boff, a.drec = bf(W.drec) + boff;
        a.dlg_off = W.dlg - boff + (long long)bb0 * 48 * 128, a.xin_off = W.xin - boff + (long long)bb0 * 32 * 128;
        a.stat = reinterpret_cast<const float*>(ws + W.stat) + (long long)bb0 * (32 * 128 * 4);
        a.stat_t_stride = (long long)L.NBBT * (32 * 128 * 4);
        a.feature = fo->feature + r0 * T * F, a.h0 = in->h0 + r0 * D;
        a.d_feature = up->d_feature + r0 * T * F;
        a.d_actions = gin->d_actions ? gin->d_actions + r0 * T * A : nullptr;
        a.d_h0 = gin->d_h0 + r0 * D, a.d_z0 = gin->d_z0 + r0 * 16;
        a.bar = reinterpret_cast<unsigned*>(ws + W.bar + 256 * (size_t)bb0);
        a.status = reinterpret_cast<int*>(ws + W.bar + 256 * (size_t)L.NBBT);
        a.timing = (g == 0 && getenv("RSSM_WIDE_TIMING")) ? reinterpret_cast<unsigned long long*>(ws + W.bar + 256 * (size_t)(L.NBBT + 1)) : nullptr;
        a.exp = getenv("RSSM_WIDE_EXP") ? atoi(getenv("RSSM_WIDE_EXP")) : 0;
        a.cs = wide_cluster_size(L.NSL);
        g_launches.fetch_add(1);
        if (check_cuda(rssm::launch_mrssm_wide_bwd(a, s), "wide mrssm backward launch")) return 1;
    }
    // ---- embedding gradients ------------------------------------------------------------------------------------------------
    {
        rssm::WideDembedArgs e{};
        e.B = d->B, e.T = T, e.D = D, e.NBBT = L.NBBT, e.dt_stride = W.dt_stride;
        e.dah = bf(W.drec) + (long long)rssm::wide::DP_AH * plane, e.dvh = bf(W.drec) + (long long)rssm::wide::DP_VH * plane;
        e.pWaeT = bf(W.pWaeT), e.pWveT = bf(W.pWveT), e.d_embed_a = gin->d_embed_a, e.d_embed_v = gin->d_embed_v;
        g_launches.fetch_add(1);
        if (check_cuda(rssm::launch_wide_dembed(e, s), "wide embedding-gradient launch")) return 1;
    }
    if (gw == nullptr) return 0;
    // ---- weight gradients: tile table ------------------------------------------------------------------------------------------
    g_launches.fetch_add(3);
    if (check_cuda(rssm::launch_wide_pack_rows(in->h0, d->B, 1, D, D, 0, bf(W.h0p), L.NBBT, s), "h0 packing launch")) return 1;
    if (check_cuda(rssm::launch_wide_pack_rows(in->embed_a, d->B, T, 64, 64, 0, bf(W.emb_a), L.NBBT, s), "embedding packing launch")) return 1;
    if (check_cuda(rssm::launch_wide_pack_rows(in->embed_v, d->B, T, 64, 64, 0, bf(W.emb_v), L.NBBT, s), "embedding packing launch")) return 1;
    static thread_local rssm::WideWgradTileTable table;  // host scratch only: passed BY VALUE to the upload kernel below
    rssm::WideWgradTile* tiles = table.t;
    int nt = 0;
    const long long bstrideD = (long long)D * 128;
    const __nv_bfloat16* dr = bf(W.drec);
    auto dplane = [&](int pl, int f0) { return dr + (long long)pl * plane + (long long)(f0 / 8) * 1024; };
    auto fplane = [&](int pl, int f0) { return rec + (long long)pl * plane + (long long)(f0 / 8) * 1024; };
    auto add = [&](const __nv_bfloat16* y, long long yt, long long yb, const __nv_bfloat16* x, const __nv_bfloat16* x0, long long xt, long long xb,
                   int shift, int N, int mvalid, int nvalid, float* dW, float* db, long long sm, long long sn) {
        rssm::WideWgradTile& t = tiles[nt++];
        t.y = y, t.x = x, t.x0 = x0, t.y_tstride = yt, t.y_bstride = yb, t.x_tstride = xt, t.x_bstride = xb, t.x_shift = shift;
        t.N = N, t.mvalid = mvalid, t.nvalid = nvalid, t.dW = dW, t.db = db, t.sm = sm, t.sn = sn;
    };
    const int MT = D / 128;
    // Y = gradient plane `ypl`, X = record plane `xpl` (all D input features, in N tiles of <= 256)
    auto dense = [&](int ypl, int xpl, int shift, float* dW, long long ldw, float* db) {
        for (int mt = 0; mt < MT; ++mt)
            for (int n0 = 0; n0 < D; n0 += 256) {
                const int N = D - n0 < 256 ? D - n0 : 256;
                add(dplane(ypl, mt * 128), W.dt_stride, bstrideD, fplane(xpl, n0), shift ? bf(W.h0p) + (long long)(n0 / 8) * 1024 : nullptr, tstride,
                    bstrideD, shift, N, 128, N, dW + (long long)mt * 128 * ldw + n0, (n0 == 0 && db) ? db + mt * 128 : nullptr, ldw, 1);
            }
    };
    using namespace rssm::wide;
    const int ih_pl[3] = {DP_GR, DP_GZ, DP_GIN}, hh_pl[3] = {DP_GR, DP_GZ, DP_GHN};
    for (int g = 0; g < 3; ++g) {
        dense(ih_pl[g], P_X2, 0, gw->w_ih + (size_t)g * D * D, D, gw->b_ih + g * D);
        dense(hh_pl[g], P_HB, 1, gw->w_hh + (size_t)g * D * D, D, gw->b_hh + g * D);
    }
    dense(DP_X2, P_HID1, 0, gw->asp_w2, D, gw->asp_b2);
    dense(DP_PH, P_HB, 0, gw->pr_w1, D, gw->pr_b1);
    dense(DP_AH, P_HB, 0, gw->au_w1, D + 64, gw->au_b1);
    dense(DP_VH, P_HB, 0, gw->vi_w1, D + 64, gw->vi_b1);
    const long long embt = (long long)L.NBBT * 64 * 128;
    for (int mt = 0; mt < MT; ++mt) {
        add(dplane(DP_AH, mt * 128), W.dt_stride, bstrideD, bf(W.emb_a), nullptr, embt, 64 * 128, 0, 64, 128, 64,
            gw->au_w1 + (size_t)mt * 128 * (D + 64) + D, nullptr, D + 64, 1);
        add(dplane(DP_VH, mt * 128), W.dt_stride, bstrideD, bf(W.emb_v), nullptr, embt, 64 * 128, 0, 64, 128, 64,
            gw->vi_w1 + (size_t)mt * 128 * (D + 64) + D, nullptr, D + 64, 1);
        // first projector layer: the operand plane is [z_{t-1} (16) | a_t (A) | 0]
        add(dplane(DP_H1, mt * 128), W.dt_stride, bstrideD, dr + W.xin, nullptr, W.dt_stride, 32 * 128, 0, 16, 128, 16,
            gw->asp_w1 + (size_t)mt * 128 * (A + 16) + A, gw->asp_b1 + mt * 128, A + 16, 1);
        add(dplane(DP_H1, mt * 128), W.dt_stride, bstrideD, dr + W.xin + 2048, nullptr, W.dt_stride, 32 * 128, 0, 16, 128, A,
            gw->asp_w1 + (size_t)mt * 128 * (A + 16), nullptr, A + 16, 1);
    }
    float* w2g[3] = {gw->pr_w2, gw->au_w2, gw->vi_w2};
    float* b2g[3] = {gw->pr_b2, gw->au_b2, gw->vi_b2};
    for (int h = 0; h < 3; ++h) {
        const __nv_bfloat16* dl = dr + W.dlg + (long long)h * 2048;
        for (int mt = 0; mt < MT; ++mt)  // dW2_h[n][m] = sum_rows hid_h[row][m] * dlogit_h[row][n]
            add(fplane(P_PH + h, mt * 128), tstride, bstrideD, dl, nullptr, W.dt_stride, 48 * 128, 0, 16, 128, 16, w2g[h] + mt * 128, nullptr, 1, D);
        add(bf(W.ones), 0, 0, dl, nullptr, W.dt_stride, 48 * 128, 0, 16, 1, 16, b2g[h], nullptr, 0, 1);  // column sums
    }
    if (nt > MAX_WIDE_TILES) return fail("internal: %d weight-gradient tiles", nt);
    table.n = nt;
    g_launches.fetch_add(1);
    if (check_cuda(rssm::launch_wide_wgrad_tiles_upload(table, reinterpret_cast<rssm::WideWgradTile*>(ws + W.tiles), s), "tile table upload launch"))
        return 1;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int nblocks = T * L.NBBT;
    int nsplit = (3 * nsm) / nt;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > nblocks) nsplit = nblocks;
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_wide_wgrad(reinterpret_cast<const rssm::WideWgradTile*>(ws + W.tiles), nt, nsplit, T, L.NBBT, bf(W.ones), s),
                      "wide weight-gradient launch");
}

}  // namespace

extern "C" {

int rssm_abi_version(void) { return RSSM_ABI_VERSION; }
const char* rssm_last_error(void) { return g_err; }
long long rssm_kernel_launch_count(void) { return g_launches.load(); }

size_t rssm_mrssm_saved_bytes(const RssmMrssmDims* d) {
    if (!d) return 0;
    if (!is_wide(d)) return (size_t)d->B * d->T * MRSSM_SAVED_FLOATS * (d->precision == RSSM_PRECISION_FP32 ? 4 : 2);
    if (check_mrssm_wide(d)) return 0;
    return wide_saved_planes_bytes(d) + (size_t)((d->B + 127) / 128) * 128 * d->T * 32 * 4;
}
size_t rssm_mrssm_workspace_bytes(const RssmMrssmDims* d, int pass) {
    if (!d || !is_wide(d) || check_mrssm_wide(d)) return 0;
    WideLayout L;
    if (wide_layout(d, true, &L)) return 0;
    return pass == 0 ? L.total_fwd : wide_bwd_workspace(d, L);
}

// ---------------------------------------------------------------------------------------------------
// MoPoE-MRSSM
// ---------------------------------------------------------------------------------------------------
static int mrssm_fwd_common(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in,
                            const RssmMrssmOutputs* out, void* stream, bool imagine) {
    if (!is_wide(d) && check_mrssm(d)) return 1;
    REQUIRE(w); REQUIRE(in); REQUIRE(out);
    REQUIRE(in->actions); REQUIRE(in->h0); REQUIRE(in->z0); REQUIRE(out->feature); REQUIRE(out->prior_probs);
    REQUIRE(w->asp_w1); REQUIRE(w->w_ih); REQUIRE(w->w_hh); REQUIRE(w->pr_w1); REQUIRE(w->pr_w2);
    if (imagine) {
        REQUIRE(in->u_prior);
    } else {
        REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->u_post); REQUIRE(out->post_probs); REQUIRE(out->kl);
        REQUIRE(w->au_w1); REQUIRE(w->au_w2); REQUIRE(w->vi_w1); REQUIRE(w->vi_w2);
    }
    if (is_wide(d)) return wide_mrssm_fwd(d, w, in, out, static_cast<cudaStream_t>(stream), imagine);
    rssm::MrssmFwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.K = d->K, a.unimodal = d->unimodal ? 1 : 0, a.w = *w;
    a.actions = in->actions, a.embed_a = in->embed_a, a.embed_v = in->embed_v, a.h0 = in->h0, a.z0 = in->z0;
    a.u_post = in->u_post, a.u_prior = in->u_prior;
    a.feature = out->feature, a.prior_probs = out->prior_probs, a.post_probs = out->post_probs;
    a.prior_stoch = out->prior_stoch, a.kl = out->kl, a.saved = imagine ? nullptr : out->saved;
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_mrssm_fwd(a, d->precision, imagine, static_cast<cudaStream_t>(stream)),
                      imagine ? "mrssm imagine launch" : "mrssm forward launch");
}

int rssm_mrssm_rollout_fwd(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in, const RssmMrssmOutputs* out,
                           void* stream) {
    return mrssm_fwd_common(d, w, in, out, stream, false);
}

int rssm_mrssm_imagine_fwd(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in, const RssmMrssmOutputs* out,
                           void* stream) {
    return mrssm_fwd_common(d, w, in, out, stream, true);
}

int rssm_mrssm_rollout_bwd(const RssmMrssmDims* d, const RssmMrssmWeights* w, const RssmMrssmInputs* in, const RssmMrssmOutputs* fo,
                           const RssmMrssmUpstream* up, const RssmMrssmInputGrads* gin, const RssmMrssmWeightGrads* gw, void* stream) {
    if (!is_wide(d) && check_mrssm(d)) return 1;
    REQUIRE(w); REQUIRE(in); REQUIRE(fo); REQUIRE(up); REQUIRE(gin);
    REQUIRE(in->actions); REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->h0); REQUIRE(in->z0);
    REQUIRE(fo->feature); REQUIRE(fo->prior_probs); REQUIRE(fo->post_probs); REQUIRE(fo->saved);
    REQUIRE(up->d_feature); REQUIRE(gin->d_embed_a); REQUIRE(gin->d_embed_v); REQUIRE(gin->d_h0); REQUIRE(gin->d_z0);
    if (is_wide(d)) return wide_mrssm_bwd(d, w, in, fo, up, gin, gw, static_cast<cudaStream_t>(stream));
    REQUIRE(gin->dpre);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rssm::MrssmBwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.K = d->K, a.unimodal = d->unimodal ? 1 : 0, a.kl_wq = up->kl_wq, a.kl_wp = up->kl_wp, a.w = *w;
    a.h0 = in->h0, a.feature = fo->feature, a.prior_probs = fo->prior_probs, a.post_probs = fo->post_probs, a.saved = fo->saved;
    a.d_feature = up->d_feature, a.d_prior_probs = up->d_prior_probs, a.d_post_probs = up->d_post_probs;
    a.d_prior_stoch = up->d_prior_stoch, a.d_kl = up->d_kl;
    a.dpre = gin->dpre, a.d_actions = gin->d_actions, a.d_embed_a = gin->d_embed_a, a.d_embed_v = gin->d_embed_v;
    a.d_h0 = gin->d_h0, a.d_z0 = gin->d_z0;
    g_launches.fetch_add(1);
    if (check_cuda(rssm::launch_mrssm_bwd(a, d->precision, s), "mrssm backward launch")) return 1;
    if (gw == nullptr) return 0;
    return rssm_mrssm_wgrad(d, in, fo, gin->dpre, gw, stream);
}

int rssm_mrssm_wgrad(const RssmMrssmDims* d, const RssmMrssmInputs* in, const RssmMrssmOutputs* fo, const void* dpre,
                     const RssmMrssmWeightGrads* gw, void* stream) {
    if (is_wide(d)) return fail("rssm_mrssm_wgrad: the wide family computes its weight gradients inside rssm_mrssm_rollout_bwd (gw != NULL)");
    if (check_mrssm(d)) return 1;
    REQUIRE(in); REQUIRE(fo); REQUIRE(dpre); REQUIRE(gw);
    REQUIRE(in->actions); REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->h0); REQUIRE(in->z0);
    REQUIRE(fo->feature); REQUIRE(fo->saved);
    // staged-row layout and part ids: kernels.h (wgl_mr); record offsets: mrssm_kernels.cu (mrs / mrd)
    using namespace rssm::wgl_mr;
    const int A = d->A, F = 48, SV = MRSSM_SAVED_FLOATS, DPF = MRSSM_DPRE_FLOATS;
    const int RE = d->precision == RSSM_PRECISION_BF16 ? 2 : 4;  // record element size
    const void* sv = fo->saved;
    const float* feat = fo->feature;
    rssm::WgradMmaArgs j{};
    j.B = d->B, j.T = d->T;
    add_seg(j, DP, dpre, DPF, DPF, RE);
    add_seg(j, XASPZ, feat + 32, F, 16, 4, 1, in->z0, 16);  // z_prev
    add_seg(j, XASPA, in->actions, A, A);                   // action (+ zero pad to 8)
    add_seg(j, XH1, rec_at(sv, 0, RE), SV, 32, RE);         // asp hidden
    add_seg(j, XX2, rec_at(sv, 32, RE), SV, 32, RE);        // x2 (GRU input)
    add_seg(j, XHP, feat, F, 32, 4, 1, in->h0, 32);         // h_prev
    add_seg(j, XA, feat, F, 32);                            // [h | embed_a]
    add_seg(j, XA + 32, in->embed_a, 64, 64);
    add_seg(j, XV, feat, F, 32);                            // [h | embed_v]
    add_seg(j, XV + 32, in->embed_v, 64, 64);
    add_seg(j, HID, rec_at(sv, 192, RE), SV, 96, RE);       // prior / audio / vision hidden
    if (j.stride != STRIDE) return fail("internal: MRSSM staged row is %d columns, expected %d", j.stride, STRIDE);
    set_out(j, O_ASP1Z, gw->asp_w1 + A, A + 16, 16, gw->asp_b1);
    set_out(j, O_ASP1A, gw->asp_w1, A + 16, A);
    set_out(j, O_ASP2, gw->asp_w2, 32, 32, gw->asp_b2);
    set_out(j, O_IHR, gw->w_ih, 32, 32, gw->b_ih);
    set_out(j, O_IHZ, gw->w_ih + 32 * 32, 32, 32, gw->b_ih + 32);
    set_out(j, O_IHN, gw->w_ih + 64 * 32, 32, 32, gw->b_ih + 64);
    set_out(j, O_HHR, gw->w_hh, 32, 32, gw->b_hh);
    set_out(j, O_HHZ, gw->w_hh + 32 * 32, 32, 32, gw->b_hh + 32);
    set_out(j, O_HHN, gw->w_hh + 64 * 32, 32, 32, gw->b_hh + 64);
    set_out(j, O_P1, gw->pr_w1, 32, 32, gw->pr_b1);
    set_out(j, O_P2, gw->pr_w2, 32, 32, gw->pr_b2);
    set_out(j, O_A1A, gw->au_w1, 96, 96, gw->au_b1);
    set_out(j, O_A1B, gw->au_w1 + 16 * 96, 96, 96, gw->au_b1 + 16);
    set_out(j, O_A2, gw->au_w2, 32, 32, gw->au_b2);
    set_out(j, O_V1A, gw->vi_w1, 96, 96, gw->vi_b1);
    set_out(j, O_V1B_L, gw->vi_w1 + 16 * 96, 96, 48, gw->vi_b1 + 16);
    set_out(j, O_V1B_R, gw->vi_w1 + 16 * 96 + 48, 96, 48);
    set_out(j, O_V2, gw->vi_w2, 32, 32, gw->vi_b2);
    for (int i = 0; i < N_OUT; ++i)
        if (j.out[i].dW == nullptr) return fail("weight-gradient pointer of part %d is NULL", i);
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_wgrad_mma(j, 1, d->precision, static_cast<cudaStream_t>(stream)), "mrssm wgrad launch");
}

// ---------------------------------------------------------------------------------------------------
// MoPoE-MMTRSSM
// ---------------------------------------------------------------------------------------------------
static int mtrssm_fwd_common(const RssmMtrssmDims* d, const RssmMtrssmWeights* w, const RssmMtrssmInputs* in,
                             const RssmMtrssmOutputs* out, void* stream, bool imagine) {
    if (check_mtrssm(d)) return 1;
    REQUIRE(w); REQUIRE(in); REQUIRE(out);
    REQUIRE(in->actions); REQUIRE(in->deter_h0); REQUIRE(in->deter_l0); REQUIRE(in->hidden_h0); REQUIRE(in->hidden_l0);
    REQUIRE(in->stoch_h0); REQUIRE(in->stoch_l0);
    REQUIRE(out->feature); REQUIRE(out->hidden_h); REQUIRE(out->hidden_l); REQUIRE(out->prior_probs_h); REQUIRE(out->prior_probs_l);
    REQUIRE(w->l_d2h_w); REQUIRE(w->l_in_w); REQUIRE(w->h_d2h_w); REQUIRE(w->h_in_w); REQUIRE(w->lp_w1); REQUIRE(w->hp_w1);
    if ((in->u_prior_l == nullptr) != (in->u_prior_h == nullptr)) return fail("u_prior_l and u_prior_h must both be given or both NULL");
    if (imagine) {
        REQUIRE(in->u_prior_l);
    } else {
        REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->u_post_l); REQUIRE(in->u_post_h);
        REQUIRE(out->post_probs_h); REQUIRE(out->post_probs_l); REQUIRE(out->kl_l); REQUIRE(out->kl_h);
        REQUIRE(w->hq_w1); REQUIRE(w->au_w1); REQUIRE(w->vi_w1);
        if ((out->prior_stoch_l == nullptr) != (out->prior_stoch_h == nullptr))
            return fail("prior_stoch_l and prior_stoch_h must both be given or both NULL");
    }
    rssm::MtrssmFwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.KL = d->KL, a.KH = d->KH;
    a.inv_tau_l = 1.f / d->l_tau, a.inv_tau_h = 1.f / d->h_tau, a.w = *w;
    a.actions = in->actions, a.embed_a = in->embed_a, a.embed_v = in->embed_v;
    a.deter_h0 = in->deter_h0, a.deter_l0 = in->deter_l0, a.hidden_h0 = in->hidden_h0, a.hidden_l0 = in->hidden_l0;
    a.stoch_h0 = in->stoch_h0, a.stoch_l0 = in->stoch_l0;
    a.u_post_l = in->u_post_l, a.u_post_h = in->u_post_h, a.u_prior_l = in->u_prior_l, a.u_prior_h = in->u_prior_h;
    a.feature = out->feature, a.hidden_h = out->hidden_h, a.hidden_l = out->hidden_l;
    a.prior_probs_h = out->prior_probs_h, a.prior_probs_l = out->prior_probs_l;
    a.post_probs_h = out->post_probs_h, a.post_probs_l = out->post_probs_l;
    a.prior_stoch_h = out->prior_stoch_h, a.prior_stoch_l = out->prior_stoch_l;
    a.kl_l = out->kl_l, a.kl_h = out->kl_h, a.saved = imagine ? nullptr : out->saved;
    a.saved_ld = mt_saved_ld(d->precision);
    a.obs_projected = imagine ? 0 : d->obs_projected;
    a.rec_tiled = d->precision == RSSM_PRECISION_BF16_FUSED ? 1 : 0;  // tile-blocked saved record (include/rssm_rollout.h)
    a.ld_feature = out->ld_feature;
    if ((out->ld_feature | out->ld_hidden | out->ld_probs | out->ld_stoch | out->ld_kl) != 0) {  // grouped rows
        if (imagine || !a.rec_tiled) return fail("grouped output rows (ld_* != 0) need the posterior rollout under RSSM_PRECISION_BF16_FUSED; pass 0 here");
        if (out->ld_feature != MTRSSM_ROW_PITCH || out->ld_hidden != MTRSSM_ROW_PITCH || out->ld_probs != MTRSSM_ROW_PITCH ||
            out->ld_stoch != MTRSSM_ROW_PITCH || out->ld_kl != 2)
            return fail("output row pitches %d %d %d %d %d: all 0 (dense) or %d %d %d %d 2 (grouped rows)", out->ld_feature, out->ld_hidden,
                        out->ld_probs, out->ld_stoch, out->ld_kl, MTRSSM_ROW_PITCH, MTRSSM_ROW_PITCH, MTRSSM_ROW_PITCH, MTRSSM_ROW_PITCH);
    }
    g_launches.fetch_add(1);
    // bf16 policies, posterior rollout: two warps per tile (mtrssm_fwd2.cu); imagination and the fp32-parity policy: one warp per tile
    if (!imagine && d->precision != RSSM_PRECISION_FP32 && getenv("RSSM_FWD_ONE_WARP") == nullptr)
        return check_cuda(rssm::launch_mtrssm_fwd2(a, static_cast<cudaStream_t>(stream)), "mtrssm forward launch");
    return check_cuda(rssm::launch_mtrssm_fwd(a, mt_kernel_precision(d->precision), imagine, static_cast<cudaStream_t>(stream)),
                      imagine ? "mtrssm imagine launch" : "mtrssm forward launch");
}

int rssm_mtrssm_rollout_fwd(const RssmMtrssmDims* d, const RssmMtrssmWeights* w, const RssmMtrssmInputs* in,
                            const RssmMtrssmOutputs* out, void* stream) {
    return mtrssm_fwd_common(d, w, in, out, stream, false);
}

int rssm_mtrssm_imagine_fwd(const RssmMtrssmDims* d, const RssmMtrssmWeights* w, const RssmMtrssmInputs* in,
                            const RssmMtrssmOutputs* out, void* stream) {
    return mtrssm_fwd_common(d, w, in, out, stream, true);
}

int rssm_mtrssm_rollout_bwd(const RssmMtrssmDims* d, const RssmMtrssmWeights* w, const RssmMtrssmInputs* in,
                            const RssmMtrssmOutputs* fo, const RssmMtrssmUpstream* up, const RssmMtrssmInputGrads* gin,
                            const RssmMtrssmWeightGrads* gw, void* stream) {
    if (check_mtrssm(d)) return 1;
    REQUIRE(w); REQUIRE(in); REQUIRE(fo); REQUIRE(up); REQUIRE(gin);
    REQUIRE(in->actions); REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->deter_h0); REQUIRE(in->deter_l0);
    REQUIRE(in->stoch_h0); REQUIRE(in->stoch_l0);
    REQUIRE(fo->feature); REQUIRE(fo->prior_probs_h); REQUIRE(fo->prior_probs_l); REQUIRE(fo->post_probs_h); REQUIRE(fo->post_probs_l);
    REQUIRE(fo->saved); REQUIRE(up->d_feature); REQUIRE(gin->d_embed_a); REQUIRE(gin->d_embed_v);
    const bool fused = gw != nullptr && d->precision == RSSM_PRECISION_BF16_FUSED;
    if (!fused) REQUIRE(gin->dpre);
    REQUIRE(gin->d_deter_h0); REQUIRE(gin->d_deter_l0); REQUIRE(gin->d_hidden_h0); REQUIRE(gin->d_hidden_l0);
    REQUIRE(gin->d_stoch_h0); REQUIRE(gin->d_stoch_l0);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rssm::MtrssmBwdArgs a{};
    a.B = d->B, a.T = d->T, a.A = d->A, a.KL = d->KL, a.KH = d->KH;
    a.inv_tau_l = 1.f / d->l_tau, a.inv_tau_h = 1.f / d->h_tau, a.kl_wq = up->kl_wq, a.kl_wp = up->kl_wp, a.w = *w;
    a.feature = fo->feature, a.prior_probs_h = fo->prior_probs_h, a.prior_probs_l = fo->prior_probs_l;
    a.post_probs_h = fo->post_probs_h, a.post_probs_l = fo->post_probs_l, a.saved = fo->saved, a.saved_ld = mt_saved_ld(d->precision);
    a.embed_a = in->embed_a, a.embed_v = in->embed_v, a.actions = in->actions;
    a.deter_h0 = in->deter_h0, a.deter_l0 = in->deter_l0, a.stoch_h0 = in->stoch_h0, a.stoch_l0 = in->stoch_l0;
    a.d_feature = up->d_feature, a.d_prior_probs_h = up->d_prior_probs_h, a.d_prior_probs_l = up->d_prior_probs_l;
    a.d_post_probs_h = up->d_post_probs_h, a.d_post_probs_l = up->d_post_probs_l;
    a.d_prior_stoch_h = up->d_prior_stoch_h, a.d_prior_stoch_l = up->d_prior_stoch_l, a.d_kl_l = up->d_kl_l, a.d_kl_h = up->d_kl_h;
    a.d_hidden_h = up->d_hidden_h, a.d_hidden_l = up->d_hidden_l;
    a.dpre = gin->dpre, a.d_actions = gin->d_actions, a.d_embed_a = gin->d_embed_a, a.d_embed_v = gin->d_embed_v;
    a.d_deter_h0 = gin->d_deter_h0, a.d_deter_l0 = gin->d_deter_l0, a.d_hidden_h0 = gin->d_hidden_h0;
    a.d_hidden_l0 = gin->d_hidden_l0, a.d_stoch_h0 = gin->d_stoch_h0, a.d_stoch_l0 = gin->d_stoch_l0;
    a.obs_projected = d->obs_projected;
    a.rec_tiled = fused ? 1 : 0;
    a.ld_feature = fo->ld_feature;
    if ((fo->ld_feature | fo->ld_probs) != 0) {
        if (!fused) return fail("grouped output rows (ld_* != 0) need the fused backward; pass 0 here");
        if (fo->ld_feature != MTRSSM_ROW_PITCH || fo->ld_probs != MTRSSM_ROW_PITCH)
            return fail("forward-output row pitches %d %d: both 0 (dense) or both %d (grouped rows)", fo->ld_feature, fo->ld_probs, MTRSSM_ROW_PITCH);
    }
    if (d->obs_projected && !fused) return fail("obs_projected needs the fused backward (gw != NULL, RSSM_PRECISION_BF16_FUSED)");
    if (d->precision == RSSM_PRECISION_BF16_FUSED && !fused)
        return fail("RSSM_PRECISION_BF16_FUSED writes the saved record tile-blocked: its backward needs the weight-gradient pointers (gw != NULL)");
    if (fused) {
        const float* const* gp = reinterpret_cast<const float* const*>(gw);
        for (size_t i = 0; i < sizeof(RssmMtrssmWeightGrads) / sizeof(float*); ++i)
            if (gp[i] == nullptr) return fail("weight-gradient pointer %d is NULL", (int)i);
        g_launches.fetch_add(1);
        return check_cuda(rssm::launch_mtrssm_bwd_fused(a, *gw, s), "mtrssm fused backward launch");
    }
    g_launches.fetch_add(1);
    if (check_cuda(rssm::launch_mtrssm_bwd(a, mt_kernel_precision(d->precision), s), "mtrssm backward launch")) return 1;
    if (gw == nullptr) return 0;
    return rssm_mtrssm_wgrad(d, in, fo, gin->dpre, gw, stream);
}

int rssm_mtrssm_wgrad(const RssmMtrssmDims* d, const RssmMtrssmInputs* in, const RssmMtrssmOutputs* fo, const void* dpre,
                      const RssmMtrssmWeightGrads* gw, void* stream) {
    if (check_mtrssm(d)) return 1;
    if (d->precision == RSSM_PRECISION_BF16_FUSED)
        return fail("rssm_mtrssm_wgrad reads the row-layout record of RSSM_PRECISION_BF16 / _FP32; RSSM_PRECISION_BF16_FUSED computes the "
                    "weight gradients inside rssm_mtrssm_rollout_bwd");
    REQUIRE(in); REQUIRE(fo); REQUIRE(dpre); REQUIRE(gw);
    REQUIRE(in->actions); REQUIRE(in->embed_a); REQUIRE(in->embed_v); REQUIRE(in->deter_h0); REQUIRE(in->deter_l0);
    REQUIRE(in->stoch_h0); REQUIRE(in->stoch_l0); REQUIRE(fo->feature); REQUIRE(fo->saved);
    // staged-row layout and part ids: kernels.h (wgl_mt); record offsets: mtrssm_kernels.cu (mts / mtd);
    // feature = [d_h 0 | z_h 32 | d_l 48 | z_l 80]
    using namespace rssm::wgl_mt;
    const int A = d->A, F = 96, DPF = MTRSSM_DPRE_FLOATS, LDIN = A + 32;
    const int SV = mt_saved_ld(d->precision);
    const int RE = d->precision == RSSM_PRECISION_FP32 ? 4 : 2;  // record element size
    const void* sv = fo->saved;
    const float* feat = fo->feature;
    rssm::WgradMmaArgs j{};
    j.B = d->B, j.T = d->T;
    add_seg(j, DP, dpre, DPF, 304, RE);
    add_seg(j, XLD, feat + 48, F, 32, 4, 1, in->deter_l0, 32);       // d_l_prev
    add_seg(j, XLZ, feat + 80, F, 16, 4, 1, in->stoch_l0, 16);       // z_l_prev
    add_seg(j, XLZ + 16, feat + 32, F, 16, 4, 1, in->stoch_h0, 16);  // z_h_prev
    add_seg(j, XLA, in->actions, A, A);                               // action (+ zero pad to 8)
    add_seg(j, XHD, feat + 0, F, 32, 4, 1, in->deter_h0, 32);        // d_h_prev
    add_seg(j, XHI, feat + 32, F, 16, 4, 1, in->stoch_h0, 16);       // z_h_prev
    add_seg(j, XQ, feat + 48, F, 32);                            // [d_l | d_h]
    add_seg(j, XQ + 32, feat + 0, F, 32);
    add_seg(j, XA, feat + 48, F, 32);                            // [d_l | embed_a]
    add_seg(j, XA + 32, in->embed_a, 64, 64);
    add_seg(j, XV, feat + 48, F, 32);                            // [d_l | embed_v]
    add_seg(j, XV + 32, in->embed_v, 64, 64);
    add_seg(j, HID, sv, SV, 160, RE);                            // the five head hiddens
    if (j.stride != STRIDE) return fail("internal: MMTRSSM staged row is %d columns, expected %d", j.stride, STRIDE);
    // l_rnn: pre_l = _d2h(d_l_prev) + _input2h([action | z_l_prev | z_h_prev]); both biases see sum(dpre_l)
    set_out(j, O_LD, gw->l_d2h_w, 32, 32, gw->l_d2h_b, gw->l_in_b);
    set_out(j, O_LIZ, gw->l_in_w + A, LDIN, 32);
    set_out(j, O_LIA, gw->l_in_w, LDIN, A);
    set_out(j, O_HD, gw->h_d2h_w, 32, 32, gw->h_d2h_b, gw->h_in_b);
    set_out(j, O_HI, gw->h_in_w, 16, 16);
    set_out(j, O_LP1, gw->lp_w1, 32, 32, gw->lp_b1);
    set_out(j, O_LP2, gw->lp_w2, 32, 32, gw->lp_b2);
    set_out(j, O_HP1, gw->hp_w1, 32, 32, gw->hp_b1);
    set_out(j, O_HP2, gw->hp_w2, 32, 32, gw->hp_b2);
    set_out(j, O_HQ1, gw->hq_w1, 64, 64, gw->hq_b1);
    set_out(j, O_HQ2, gw->hq_w2, 32, 32, gw->hq_b2);
    set_out(j, O_A1A, gw->au_w1, 96, 96, gw->au_b1);
    set_out(j, O_A1B, gw->au_w1 + 16 * 96, 96, 96, gw->au_b1 + 16);
    set_out(j, O_A2, gw->au_w2, 32, 32, gw->au_b2);
    set_out(j, O_V1A, gw->vi_w1, 96, 96, gw->vi_b1);
    set_out(j, O_V1B, gw->vi_w1 + 16 * 96, 96, 96, gw->vi_b1 + 16);
    set_out(j, O_V2, gw->vi_w2, 32, 32, gw->vi_b2);
    for (int i = 0; i < N_OUT; ++i)
        if (j.out[i].dW == nullptr) return fail("weight-gradient pointer of part %d is NULL", i);
    g_launches.fetch_add(1);
    // bf16 records with the padded row lengths: slab-staged kernel (one bulk copy per source tensor and 32-row block)
    const bool slab_ok = d->precision == RSSM_PRECISION_BF16 && aligned16(dpre) && aligned16(sv) && aligned16(feat) &&
                         aligned16(in->embed_a) && aligned16(in->embed_v) && aligned16(in->deter_l0) && aligned16(in->deter_h0) &&
                         aligned16(in->stoch_l0) && aligned16(in->stoch_h0) && getenv("RSSM_WGRAD_GENERIC") == nullptr;
    if (slab_ok) {
        rssm::WgradMtSlabArgs k{};
        k.B = d->B, k.T = d->T, k.A = A;
        k.dpre = static_cast<const __nv_bfloat16*>(dpre), k.saved = static_cast<const __nv_bfloat16*>(sv);
        k.feature = feat, k.embed_a = in->embed_a, k.embed_v = in->embed_v, k.actions = in->actions;
        k.deter_l0 = in->deter_l0, k.deter_h0 = in->deter_h0, k.stoch_l0 = in->stoch_l0, k.stoch_h0 = in->stoch_h0;
        for (int i = 0; i < N_OUT; ++i) k.out[i] = j.out[i];
        return check_cuda(rssm::launch_wgrad_mt_slab(k, static_cast<cudaStream_t>(stream)), "mtrssm wgrad (slab) launch");
    }
    return check_cuda(rssm::launch_wgrad_mma(j, 0, mt_kernel_precision(d->precision), static_cast<cudaStream_t>(stream)), "mtrssm wgrad launch");
}

// ---------------------------------------------------------------------------------------------------
// reconstruction likelihood (objective.py:7-23)
// ---------------------------------------------------------------------------------------------------
size_t rssm_gaussian_nll_workspace_bytes(void) {
    return 64 + sizeof(double) * RSSM_NLL_MAX_SEGMENTS * (size_t)rssm::nll_ctas_per_segment(1);
}

static int nll_args(const RssmNllPair* pairs, int n_pairs, int pred_dtype, bool backward, rssm::NllArgs* a) {
    REQUIRE(pairs);
    if (n_pairs < 1 || n_pairs > RSSM_NLL_MAX_SEGMENTS) return fail("n_pairs must be 1..%d (got %d)", RSSM_NLL_MAX_SEGMENTS, n_pairs);
    if (pred_dtype != RSSM_DTYPE_F32 && pred_dtype != RSSM_DTYPE_BF16 && pred_dtype != RSSM_DTYPE_F16) return fail("bad pred_dtype %d", pred_dtype);
    a->nseg = n_pairs, a->pred_dtype = pred_dtype;
    for (int i = 0; i < n_pairs; ++i) {
        const RssmNllPair& p = pairs[i];
        REQUIRE(p.prediction); REQUIRE(p.target);
        if (p.n_elems < 1 || p.n_batch < 1 || p.n_elems % p.n_batch) return fail("pair %d: n_elems=%zu must be a positive multiple of n_batch=%zu", i, p.n_elems, p.n_batch);
        if (!(p.scale > 0.f)) return fail("pair %d: scale must be positive (got %g)", i, (double)p.scale);
        if (!aligned16(p.prediction) || !aligned16(p.target)) return fail("pair %d: prediction / target must be 16-byte aligned", i);
        rssm::NllSeg& s = a->seg[i];
        s.prediction = p.prediction, s.target = p.target, s.n = p.n_elems;
        s.sq_coeff = 0.5 / ((double)p.scale * (double)p.scale * (double)p.n_batch);
        s.constant = (double)(p.n_elems / p.n_batch) * (log((double)p.scale) + 0.9189385332046727418 /* 0.5 log(2 pi) */);
        if (backward) {
            REQUIRE(p.d_prediction);
            if (!aligned16(p.d_prediction) || !aligned16(p.d_target)) return fail("pair %d: gradient outputs must be 16-byte aligned", i);
            s.d_loss = p.d_loss, s.d_prediction = p.d_prediction, s.d_target = p.d_target;
        } else {
            REQUIRE(p.loss);
            s.loss = p.loss;
        }
    }
    return 0;
}

int rssm_gaussian_nll_fwd(const RssmNllPair* pairs, int n_pairs, int pred_dtype, void* workspace, size_t workspace_bytes, void* stream) {
    rssm::NllArgs a{};
    if (nll_args(pairs, n_pairs, pred_dtype, false, &a)) return 1;
    REQUIRE(workspace);
    if (workspace_bytes < rssm_gaussian_nll_workspace_bytes() || !aligned16(workspace))
        return fail("workspace: need %zu zero-filled, 16-byte aligned bytes (got %zu)", rssm_gaussian_nll_workspace_bytes(), workspace_bytes);
    a.tickets = static_cast<unsigned*>(workspace);
    a.partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 64);
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_gaussian_nll(a, false, static_cast<cudaStream_t>(stream)), "gaussian nll forward launch");
}

int rssm_gaussian_nll_bwd(const RssmNllPair* pairs, int n_pairs, int pred_dtype, void* stream) {
    rssm::NllArgs a{};
    if (nll_args(pairs, n_pairs, pred_dtype, true, &a)) return 1;
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_gaussian_nll(a, true, static_cast<cudaStream_t>(stream)), "gaussian nll backward launch");
}


// ---- one-shot peer-memory allreduce of the gradient bucket ------------------------------------------------------------------
// region layout: [flags 2 x 8 u32 = 64 B | status word @64 | pad to 256 B | bucket 0 | bucket 1], buckets 256-byte aligned
static size_t p2p_bucket_stride(size_t n) { return (n * sizeof(float) + 255) / 256 * 256; }
size_t rssm_p2p_region_bytes(size_t n) { return 256 + 2 * p2p_bucket_stride(n); }
float* rssm_p2p_bucket(void* region, size_t n, int slot) {
    return reinterpret_cast<float*>(static_cast<char*>(region) + 256 + (size_t)(slot & 1) * p2p_bucket_stride(n));
}
int rssm_p2p_alloc(size_t bytes, void** region) {
    REQUIRE(region);
    if (bytes < 256) return fail("p2p region too small (%zu bytes): size it with rssm_p2p_region_bytes", bytes);
    if (check_cuda(cudaMalloc(region, bytes), "p2p region cudaMalloc")) return 1;
    return check_cuda(cudaMemset(*region, 0, bytes), "p2p region memset");
}
int rssm_p2p_free(void* region) { return check_cuda(cudaFree(region), "p2p region cudaFree"); }
int rssm_p2p_export(void* region, unsigned char handle[64]) {
    REQUIRE(region); REQUIRE(handle);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    if (check_cuda(cudaIpcGetMemHandle(&h, region), "cudaIpcGetMemHandle")) return 1;
    memcpy(handle, &h, 64);
    return 0;
}
int rssm_p2p_import(const unsigned char handle[64], void** peer_region) {
    REQUIRE(handle); REQUIRE(peer_region);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    return check_cuda(cudaIpcOpenMemHandle(peer_region, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer access between the GPUs of this box)");
}
int rssm_p2p_close(void* peer_region) { return check_cuda(cudaIpcCloseMemHandle(peer_region), "cudaIpcCloseMemHandle"); }
static int p2p_check(const RssmP2pComm* c) {
    REQUIRE(c);
    if (c->world < 1 || c->world > RSSM_P2P_MAX_RANKS || c->rank < 0 || c->rank >= c->world)
        return fail("p2p comm: world %d / rank %d (at most %d ranks of one box)", c->world, c->rank, RSSM_P2P_MAX_RANKS);
    for (int r = 0; r < c->world; ++r)
        if (c->regions[r] == nullptr) return fail("p2p comm: region of rank %d is NULL", r);
    if (c->n < 1) return fail("p2p comm: empty bucket");
    return 0;
}
int rssm_p2p_allreduce_mean(const RssmP2pComm* c, long long step, const float* src, float* out, int timeout_ms, void* stream) {
    if (p2p_check(c)) return 1;
    REQUIRE(out);
    if (step < 0) return fail("p2p allreduce: step must count up from 0 (got %lld)", step);
    if (!aligned16(out)) return fail("p2p allreduce: out must be 16-byte aligned");
    rssm::P2pAllreduceArgs a{};
    a.world = c->world, a.rank = c->rank, a.n = c->n, a.slot = (int)(step & 1), a.out = out;
    a.epoch = (uint32_t)(step / 2 + 1);  // monotonic per slot, never 0 (the flags start zero-filled)
    for (int r = 0; r < c->world; ++r) {
        a.data[r] = rssm_p2p_bucket(c->regions[r], c->n, a.slot);
        a.flags[r] = static_cast<uint32_t*>(c->regions[r]);
    }
    a.status = static_cast<uint32_t*>(c->regions[c->rank]) + 16;
    if (src != nullptr &&
        check_cuda(cudaMemcpyAsync(rssm_p2p_bucket(c->regions[c->rank], c->n, a.slot), src, c->n * sizeof(float), cudaMemcpyDeviceToDevice,
                                   static_cast<cudaStream_t>(stream)), "p2p allreduce: copy into the peer-mapped bucket"))
        return 1;
    // clock64 ticks per millisecond: a generous constant (2.5 GHz) -- cudaDevAttrClockRate is a slow driver query (it put 0.6 ms of
    // host time into every step when asked per call), and the bound only has to be "long enough, but not forever"
    a.timeout_cycles = (long long)(timeout_ms > 0 ? timeout_ms : 2000) * 2500000LL;
    g_launches.fetch_add(1);
    return check_cuda(rssm::launch_p2p_allreduce_mean(a, static_cast<cudaStream_t>(stream)), "p2p allreduce launch");
}
int rssm_p2p_status(const RssmP2pComm* c) {
    if (p2p_check(c)) return -1;
    uint32_t v = 0;
    if (check_cuda(cudaMemcpy(&v, static_cast<uint32_t*>(c->regions[c->rank]) + 16, 4, cudaMemcpyDeviceToHost), "p2p status read")) return -1;
    if (v) fail("a p2p allreduce timed out waiting for its peers");
    return (int)v;
}

}  // extern "C"
